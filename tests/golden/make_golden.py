"""Regenerates the committed golden fixtures under tests/golden/.

Run in the authoring container only (needs /root/reference and a built
oracle/_ref):   python tests/golden/make_golden.py

Two kinds of fixtures are produced:

1. ref_tests.json -- the literal golden vectors and known answers held by the
   reference's OWN tests for this path, re-expressed as data (parsed from the
   headers, not copied as code):
     src/tests/data_tests.hpp:40-47   7x30 genotype matrix of small_modelspace.bed
     src/tests/data_tests.hpp:130-140 10x5 matrix (-1 = missing) of plinktest.bed
     src/tests/data_tests.hpp:171-190 missing positions / cumulative priors
     src/tests/model_tests.hpp:68-89  log marginal likelihood constants
     README.markdown:325-327          var y / var x / mean x of the bundled data
   plus copies of the four tiny binary/text data fixtures those tests read
   (tests/golden/data/: plinktest.{bed,fam}, small_modelspace.{bed,fam}).

2. ref_outputs.npz -- outputs of the UNMODIFIED reference (oracle/_ref) on those
   fixtures and on a seeded synthetic data set (bmagwa_b200.synth): decoded
   columns, moment cache, missing index, scan probabilities for several model
   states, log-likelihood traces of add/remove sequences, proposal-tree draws,
   RNG draws, Cholesky kit results, and short fixed-seed chains.
"""
import json
import os
import re
import shutil
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

from bmagwa_b200 import synth  # noqa: E402
from oracle import ref  # noqa: E402

REF = "/root/reference"
RT = os.path.join(REF, "src", "tests")

SMALL_INI = """[datafiles]
file_fam = {d}/small_modelspace.fam
file_g = {d}/small_modelspace.bed
recode_g_to_minor_allele_count = {recode}
[sizes]
n = 30
m_g = 7
m_e = 0
[sampler]
type = PMV
do_n_iter = {do_n_iter}
n_rao = {n_rao}
n_rao_burnin = {n_rao_burnin}
adaptation = 0
delay_rejection = {delay_rejection}
verbosity = 0
thin = {thin}
save_beta = 0
n_sample_tau2_and_missing = 10
max_move_size = 7
max_SNP_neighborhood_size = 3
[thread]
n_threads = 1
basename = {out}/chain
seeds = {seed}
[model]
types = A
[prior]
e_qg = 2
var_qg = 2
use_individual_tau2 = {indiv}
R2mode_sigma2 = 0.20
nu_sigma2 = 1
nu_tau2_A = 3
s2_tau2_A = 0.02
mu_alpha = 1.0
inv_tau2_e_const_val = 0
inv_tau2_e_val = 1
"""

PLINK_INI = """[datafiles]
file_fam = {d}/plinktest.fam
file_g = {d}/plinktest.bed
recode_g_to_minor_allele_count = 0
[sizes]
n = 5
m_g = 10
m_e = 0
[sampler]
type = PMV
do_n_iter = 100
n_rao = 50
n_rao_burnin = 1
verbosity = 0
thin = 10
n_sample_tau2_and_missing = 10
max_move_size = 5
max_SNP_neighborhood_size = 2
[thread]
n_threads = 1
basename = {out}/chain
seeds = 1245
[model]
types = A
[prior]
e_qg = 2
var_qg = 4
use_individual_tau2 = 0
s2_sigma2 = 1
nu_sigma2 = 4
nu_tau2_A = 4
s2_tau2_A = 1
mu_alpha = 2.0
inv_tau2_e_const_val = 0
inv_tau2_e_val = 1
"""


def parse_int_matrix(text, name):
    m = re.search(name + r"\[\d+\]\[\d+\]\s*=\s*\{(.*?)\};", text, re.S)
    rows = re.findall(r"\{([^{}]*)\}", m.group(1))
    return [[int(v) for v in r.split(",")] for r in rows]


def parse_double_array(text, name):
    m = re.search(name + r"\[\d+\]\s*=\s*\{(.*?)\};", text, re.S)
    return [float(v) for v in m.group(1).replace("\n", " ").split(",") if v.strip()]


def golden_from_reference_tests():
    data_tests = open(os.path.join(RT, "data_tests.hpp")).read()
    first, second = data_tests.split("LoadAndHandlePlinkDataWithMissing")
    out = {
        "small_modelspace": {
            "n": 30, "m_g": 7,
            "genotypes_snp_major": parse_int_matrix(first, "genotypes"),
            "y": parse_double_array(first, "y"),
            "source": "src/tests/data_tests.hpp:40-47 (+ y literal)",
        },
        "plinktest": {
            "n": 5, "m_g": 10,
            "genotypes_snp_major": parse_int_matrix(second, "genotypes"),
            "missing": {"0": {"idx": [3], "prior_cum": [1, 2, 4]}, "3": {"idx": [1], "prior_cum": [2, 3, 4]}},
            "source": "src/tests/data_tests.hpp:130-140,171-190",
        },
        # src/tests/model_tests.hpp:36-89 (plinktest, types A+AH, nu_sigma=4, s2_sigma=1, nu_tau2=4, s2_tau2=1,
        # alpha=2 -> tau2 = 2, inv_tau2_alpha2 = 1/(4*2)); first KAT: SNP 1 added as type A.
        "model_kat": {
            "nu_sigma": 4.0, "s2_sigma": 1.0, "tau2": 2.0, "alpha": 2.0, "n": 5,
            "add_snp1_A": {"Syy": 0.58 - 0.51670588235294123702,
                           "log_det_term": "0.5*log(tau2) + 0.5*log(4*10 + 5/tau2)"},
            "add_snp1_A_then_snp5_AH": {"Syy": 0.58 - 0.56537856592135671274, "det": 810.62499999999954525265},
            "source": "src/tests/model_tests.hpp:68-89",
        },
        "readme_prints": {"var_y": 0.990225, "var_x": 0.353579, "mean_x": 0.530252,
                          "source": "README.markdown:325-327 (6 significant digits as printed)"},
    }
    return out


def run_small(tmp, out, d):
    """Reference outputs on small_modelspace (n=30, m_g=7, no missing)."""
    for indiv in (0, 1):
        ini = os.path.join(tmp, "small%d.ini" % indiv)
        open(ini, "w").write(SMALL_INI.format(d=d, out=tmp, recode=1, do_n_iter=10, n_rao=10, n_rao_burnin=1,
                                              delay_rejection=7, thin=10, seed=1245, indiv=indiv))
        R = ref.Ref(ini)
        tag = "small%d_" % indiv
        if indiv == 0:
            out["small_stats"] = np.array(list(R.data_stats().values()))
            out["small_cols_A"] = np.stack([R.get_column(j, 0) for j in range(7)], axis=1)
            out["small_moments"] = R.moments()
            out["small_prior"] = np.array(list(R.prior_params().values()))
        # RB closed-form state of src/tests/raoblackwellizer_tests.hpp:31-93: empty model, residual = y
        R.model_compute_loglik()
        R.model_set_beta_sigma2(np.array([0.1]), 0.7)
        out[tag + "scan_empty_resid_y"] = R.scan(np.zeros(30))
        # a model with two SNPs
        R.model_add(2, 11.0)
        R.model_add(5, 7.5)
        beta = np.array([0.05, 0.4, -0.3])
        R.model_set_beta_sigma2(beta, 0.9)
        pves, y_hat = R.model_compute_pve()
        out[tag + "scan_two_snps"] = R.scan(y_hat)
        out[tag + "two_snps_pves"] = pves
        out[tag + "two_snps_yhat"] = y_hat
        R.close()


def run_loglik_trace(tmp, out, d):
    """add/remove sequence on small_modelspace with explicit taus -> log-likelihood trace + final state."""
    ini = os.path.join(tmp, "small_ll.ini")
    open(ini, "w").write(SMALL_INI.format(d=d, out=tmp, recode=1, do_n_iter=10, n_rao=10, n_rao_burnin=1,
                                          delay_rejection=7, thin=10, seed=1245, indiv=0))
    R = ref.Ref(ini)
    ops = [("add", 3, 2.0), ("add", 0, 5.0), ("add", 6, 1.5), ("rem", 1, 0), ("add", 4, 9.0), ("rem", 0, 0),
           ("add", 1, 0.7), ("rem", 2, 0), ("add", 2, 3.3), ("add", 5, 4.4), ("rem", 1, 0), ("rem", 2, 0)]
    trace = [R.model_loglik()]
    for op, a, b in ops:
        if op == "add":
            R.model_add(a, b)
        else:
            R.model_remove(a)
        trace.append(R.model_loglik())
    out["ll_ops"] = np.array([[0 if o == "add" else 1, a, b] for o, a, b in ops], dtype=np.float64)
    out["ll_trace"] = np.array(trace)
    out["ll_final_xx"] = R.model_get("xx")
    out["ll_final_l"] = R.model_get("l")
    out["ll_final_xy"] = R.model_get("xy")
    out["ll_final_v"] = R.model_get("v")
    out["ll_final_loci"] = R.model_loci().astype(np.int64)
    R.model_compute_loglik()
    out["ll_final_full"] = np.array([R.model_loglik()])
    R.close()


def run_plinktest(tmp, out, d):
    ini = os.path.join(tmp, "plink.ini")
    open(ini, "w").write(PLINK_INI.format(d=d, out=tmp))
    R = ref.Ref(ini)
    out["plink_cols_raw"] = np.stack([np.stack([R.get_column(j, t) for j in range(10)], axis=1) for t in range(4)])
    out["plink_moments"] = R.moments()
    out["plink_stats"] = np.array(list(R.data_stats().values()))
    # overlay: impute SNP0/ind3 = 2, SNP3/ind1 = 1 (src/tests/data_model_tests.hpp:58-126 style)
    R.set_miss_val(0, 0, 2)
    R.set_miss_val(3, 0, 1)
    out["plink_cols_overlay"] = np.stack([np.stack([R.get_column(j, t, overlay=True) for j in range(10)], axis=1)
                                          for t in range(4)])
    mom = R.moments()
    out["plink_moments_patched"] = np.stack([R.update_prexx_cov(j, mom[j]) for j in range(10)])
    # model_tests.hpp:36-76 first KAT: add SNP 1 (type A) with inv_tau2_alpha2 = 1/(alpha^2 tau2) = 1/8
    R.model_add(1, 1.0 / 8.0)
    out["plink_ll_add1"] = np.array([R.model_loglik()])
    R.close()


def run_synthetic(tmp, out):
    n, m = 203, 300
    for indiv in (0, 1):
        ds = synth.write_dataset(tmp, "syn%d" % indiv, n=n, m_g=m, m_e=2, seed=7, miss_rate=0.01, e_qg=5, var_qg=20,
                                 use_individual_tau2=indiv)
        R = ref.Ref(ds["ini"])
        tag = "syn%d_" % indiv
        if indiv == 0:
            out["syn_stats"] = np.array(list(R.data_stats().values()))
            out["syn_moments"] = R.moments()
            out["syn_cols_A_first16"] = np.stack([R.get_column(j, 0) for j in range(16)], axis=1)
            out["syn_prior"] = np.array(list(R.prior_params().values()))
        rs = np.random.default_rng(3)
        miss = []
        for j in range(m):
            idx, _ = R.missing(j)
            for k in range(len(idx)):
                v = int(rs.integers(0, 3))
                R.set_miss_val(j, k, v)
                miss.append(v)
        out[tag + "miss_val"] = np.array(miss, dtype=np.int8)
        loci, taus = [5, 120, 299], [3.0, 4.5, 2.2]
        for s, t in zip(loci, taus):
            R.model_add(s, t)
        beta = rs.normal(size=R.model_cols()) * 0.3
        R.model_set_beta_sigma2(beta, 0.8)
        pves, y_hat = R.model_compute_pve()
        out[tag + "beta"] = beta
        out[tag + "pves"] = pves
        out[tag + "y_hat"] = y_hat
        out[tag + "xx"] = R.model_get("xx")
        out[tag + "xy"] = R.model_get("xy")
        out[tag + "loglik"] = np.array([R.model_loglik()])
        out[tag + "p_r"] = R.scan(y_hat)
        R.close()


def run_rng(out):
    g = ref.RefRng(1234, 1001.0)
    seq = []
    for i in range(64):
        seq += [g.u01(), g.normal(), g.sinvchi2_fixed(0.7), g.sinvchi2(5.0, 0.05), g.sinvchi2(0.6, 1.3), g.u01(),
                g.normal(), g.sinvchi2(2.0, 1.0)]
    out["rng_seed1234_nu1001"] = np.array(seq)


def run_dd(out):
    rs = np.random.default_rng(11)
    for m in (1, 2, 5, 10, 31, 257):
        w = rs.uniform(0.01, 1.0, size=m)
        dd = ref.RefDD(w, 99)
        twin = ref.RefRng(99, 1.0)
        rec = []
        zeroed = np.zeros(m, dtype=bool)
        for step in range(200):
            a = rs.integers(0, 4)
            i = int(rs.integers(0, m))
            if a == 0 and not zeroed[i] and (~zeroed).sum() > 1:
                dd.zero(i); zeroed[i] = True
                rec.append([0, i, dd.total(), 0])
            elif a == 1 and zeroed[i]:
                dd.unzero(i); zeroed[i] = False
                rec.append([1, i, dd.total(), 0])
            else:
                u = twin.u01()
                rec.append([2, dd.sample(), dd.total(), u])
        out["dd_w_%d" % m] = w
        out["dd_rec_%d" % m] = np.array(rec)
    # structural golden of src/tests/discrete_distribution_tests.hpp: in-order sequence for m = 10
    out["dd_inorder_10"] = np.array([7, 3, 8, 1, 9, 4, 0, 5, 2, 6])


def run_chol(out):
    rs = np.random.default_rng(5)
    k = 9
    A = rs.normal(size=(40, k))
    S = A.T @ A + np.eye(k)
    ok, U = ref.chol(S)
    out["chol_S"] = S
    out["chol_U"] = U
    out["chol_del3"] = ref.chol_downdate(U, 3)
    out["chol_del0"] = ref.chol_downdate(U, 0)
    v = rs.normal(size=k)
    U2, v2 = ref.chol_swapadj(U, 4, v)
    out["chol_swap4_U"] = U2
    out["chol_swap4_v_in"] = v
    out["chol_swap4_v"] = v2


def read_chain(base):
    def rd(name, dt):
        return np.fromfile(base + "_" + name + ".dat", dtype=dt)
    return dict(jumpdistance=rd("jumpdistance", np.uint8), move_type=rd("move_type", np.uint8),
                move_size=rd("move_size", np.uint8), modelsize=rd("modelsize", np.uint32), loci=rd("loci", np.uint32),
                log_likelihood=rd("log_likelihood", np.float64), log_prior=rd("log_prior", np.float64),
                sigma2=rd("sigma2", np.float64), pve=rd("pve", np.float64), alpha=rd("alpha", np.float64),
                rao=rd("rao", np.float64))


def run_chains(tmp, out, d):
    """Short fixed-seed chains of the unmodified reference: the accepted-move sequence goldens."""
    # (a) small_modelspace, shared tau, with delayed rejection
    for tag, indiv, dr in (("chainA", 0, 7), ("chainB", 1, 0)):
        sub = os.path.join(tmp, tag)
        os.makedirs(sub, exist_ok=True)
        ini = os.path.join(sub, "c.ini")
        open(ini, "w").write(SMALL_INI.format(d=d, out=sub, recode=1, do_n_iter=3000, n_rao=50, n_rao_burnin=20,
                                              delay_rejection=dr, thin=10, seed=4321, indiv=indiv))
        R = ref.Ref(ini)
        R.run_chain()
        R.close()
        for k, v in read_chain(os.path.join(sub, "chain0")).items():
            out["%s_%s" % (tag, k)] = v
    # (b) synthetic n=203 x m=300 with covariates (no missing), individual tau, DR on
    sub = os.path.join(tmp, "chainC")
    ds = synth.write_dataset(sub, "syn", n=203, m_g=300, m_e=2, seed=9, miss_rate=0.0, e_qg=5, var_qg=20,
                             use_individual_tau2=1, do_n_iter=4000, n_rao=100, n_rao_burnin=20, delay_rejection=10,
                             outbase=os.path.join(sub, "chain"), seeds="1234")
    R = ref.Ref(ds["ini"])
    R.run_chain()
    R.close()
    for k, v in read_chain(os.path.join(sub, "chain0")).items():
        out["chainC_%s" % k] = v


# data sets with missing genotype calls: the chain exercises DataModel::sample_missing before
# every scan, sample_missing_single at every proposed addition and the Gibbs step Sampler::sample_missing
from missing_chains import MISSING_CHAINS  # noqa: E402


def run_missing_chains(tmp, out):
    for tag, kw in MISSING_CHAINS.items():
        sub = os.path.join(tmp, tag)
        ds = synth.write_dataset(sub, "syn", outbase=os.path.join(sub, "chain"), **kw)
        R = ref.Ref(ds["ini"])
        R.run_chain()
        R.close()
        for k, v in read_chain(os.path.join(sub, "chain0")).items():
            out["%s_%s" % (tag, k)] = v


def write_chain_files_for_mcmcpos(directory):
    """_loci.dat / _modelsize.dat of two committed golden chains (chainC: 300 SNPs without missing calls, chainD: 300
    SNPs with), as chain0 / chain1 of one run."""
    g = np.load(os.path.join(HERE, "ref_outputs.npz"))
    g2 = np.load(os.path.join(HERE, "ref_missing_chains.npz"))
    for i, (src, tag) in enumerate(((g, "chainC"), (g2, "chainD"))):
        src[tag + "_loci"].astype(np.uint32).tofile(os.path.join(directory, "chain%d_loci.dat" % i))
        src[tag + "_modelsize"].astype(np.uint32).tofile(os.path.join(directory, "chain%d_modelsize.dat" % i))


def run_mcmcpos():
    """The reference's own post-processing script (bmagwa_postprocess.py mcmcpos) on those files: its text output is
    the golden for bmagwa_b200.postprocess."""
    import subprocess
    tmp = tempfile.mkdtemp()
    write_chain_files_for_mcmcpos(tmp)
    args = dict(nsnps=300, burnin=50, thin=2)
    subprocess.check_call([sys.executable, os.path.join(REF, "bmagwa_postprocess.py"), "mcmcpos", os.path.join(tmp, "chain"),
                           str(args["nsnps"]), str(args["burnin"]), str(args["thin"])], stdout=subprocess.DEVNULL)
    out = {"args": args}
    for name in ("chain0_mcmcpos.txt", "chain1_mcmcpos.txt", "chain_mcmcpos.txt"):
        out[name] = open(os.path.join(tmp, name)).read()
    json.dump(out, open(os.path.join(HERE, "ref_mcmcpos.json"), "w"))
    print("wrote ref_mcmcpos.json")


def run_typed_scan():
    """The reference's scan for several effect-type configurations (tests/typed_scan_cases.py) -> ref_typed_scan.npz."""
    from tests import typed_scan_cases as tc
    out = {}
    for tag, types, miss, indiv in tc.CASES:
        c = tc.reference_case(ref, tempfile.mkdtemp(), types, miss, indiv)
        for k in tc.GOLDEN_KEYS:
            out["%s_%s" % (tag, k)] = np.asarray(c[k])
    np.savez_compressed(os.path.join(HERE, "ref_typed_scan.npz"), **out)
    print("wrote ref_typed_scan.npz:", len(out), "arrays;", os.path.getsize(os.path.join(HERE, "ref_typed_scan.npz")), "bytes")


def main():
    if "--typed-scan-only" in sys.argv:
        assert ref.available(), "build oracle/_ref first (make -C oracle ref)"
        run_typed_scan()
        return
    if "--mcmcpos-only" in sys.argv:
        run_mcmcpos()
        return
    if "--missing-only" in sys.argv:   # only ref_missing_chains.npz (the other fixtures stay as committed)
        assert ref.available(), "build oracle/_ref first (make -C oracle ref)"
        out = {}
        run_missing_chains(tempfile.mkdtemp(), out)
        np.savez_compressed(os.path.join(HERE, "ref_missing_chains.npz"), **out)
        print("wrote", len(out), "arrays;", os.path.getsize(os.path.join(HERE, "ref_missing_chains.npz")), "bytes")
        return
    assert ref.available(), "build oracle/_ref first (make -C oracle ref)"
    d = os.path.join(HERE, "data")
    os.makedirs(d, exist_ok=True)
    for f in ("plinktest.bed", "plinktest.fam", "small_modelspace.bed", "small_modelspace.fam"):
        shutil.copyfile(os.path.join(RT, "testdata", f), os.path.join(d, f))
        os.chmod(os.path.join(d, f), 0o644)
    json.dump(golden_from_reference_tests(), open(os.path.join(HERE, "ref_tests.json"), "w"), indent=1)
    out = {}
    tmp = tempfile.mkdtemp()
    run_small(tmp, out, d)
    run_loglik_trace(tmp, out, d)
    run_plinktest(tmp, out, d)
    run_synthetic(tmp, out)
    run_rng(out)
    run_dd(out)
    run_chol(out)
    run_chains(tmp, out, d)
    np.savez_compressed(os.path.join(HERE, "ref_outputs.npz"), **out)
    out2 = {}
    run_missing_chains(tmp, out2)
    np.savez_compressed(os.path.join(HERE, "ref_missing_chains.npz"), **out2)
    run_mcmcpos()
    run_typed_scan()
    print("wrote", len(out), "arrays;", os.path.getsize(os.path.join(HERE, "ref_outputs.npz")), "bytes")


if __name__ == "__main__":
    main()
