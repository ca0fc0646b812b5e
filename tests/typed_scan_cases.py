"""Shared set-up of the multi-type scan tests: a seeded data set, a model with SNPs of
several effect types and the inputs of the scan, taken from the unmodified reference (oracle/_ref) when it is available
(CPU pinning test, golden generation) or from the committed golden file (GPU test)."""
import numpy as np

from bmagwa_b200 import synth
from oracle import cpu

NAMES = {"A": 0, "H": 1, "D": 2, "R": 3, "AH": 4}

# (tag, model.types, miss_rate, use_individual_tau2)
CASES = [
    ("H", "H", 0.0, 0), ("D", "D", 0.03, 1), ("R", "R", 0.0, 1),
    ("A_H", "A,H", 0.0, 0), ("A_H_D_R", "A,H,D,R", 0.02, 1), ("AH", "AH", 0.0, 1),
    ("A_AH", "A,AH", 0.03, 0), ("all5", "A,H,D,R,AH", 0.02, 1), ("H_A", "H,A", 0.0, 1),
]
N, M_G, M_E = 180, 90, 1


def dataset(directory, types, miss_rate, indiv):
    return synth.write_dataset(directory, "syn", n=N, m_g=M_G, m_e=M_E, seed=31, miss_rate=miss_rate, e_qg=5, var_qg=20,
                               use_individual_tau2=indiv, do_n_iter=100, n_rao=50, n_rao_burnin=1, types=types,
                               outbase=directory + "/chain", seeds="4242")


def model_for(types_codes, rng):
    """Five in-model SNPs cycling through the configured types: (snp, index into model.types, betas, taus)."""
    snps = rng.choice(M_G, size=5, replace=False)
    out = []
    for i, j in enumerate(snps):
        ti = i % len(types_codes)
        two = types_codes[ti] == 4
        out.append((int(j), ti, list(rng.normal(size=2 if two else 1) * 0.3), list(0.5 + rng.random(size=2 if two else 1))))
    return out


def columns_from_payload(payload, miss_vals_by_snp):
    """(columns0, columns, bed): additive columns with missing = 0 and with the imputed values, minor-allele recoded."""
    bed = payload.copy()
    cpu.recode_minor(bed, N, M_G)
    off, idx, _ = cpu.missing_index(bed, N, M_G)
    cols0, cols = [], []
    for j in range(M_G):
        mi = idx[off[j]:off[j + 1]]
        cols0.append(cpu.decode_column_overlay(bed, N, j, 0, mi, np.zeros(mi.size, dtype=np.int8)))
        cols.append(cpu.decode_column_overlay(bed, N, j, 0, mi, miss_vals_by_snp[j]))
    return cols0, cols, bed, off, idx


def reference_case(ref_lib, directory, types, miss_rate, indiv):
    """Runs the reference's scan on a model holding SNPs of the configured types; returns the inputs the restatement and
    the CUDA path need (everything that came out of the reference or of the seeded generator) and the reference's outputs."""
    ds = dataset(directory, types, miss_rate, indiv)
    codes = sorted(NAMES[t] for t in types.split(","))   # the reference sorts model.types (options.hpp:254)
    R = ref_lib.Ref(ds["ini"])
    rng = np.random.default_rng(5)
    miss_vals = []
    for j in range(M_G):
        cnt = R.missing(j)[0].size
        v = rng.integers(0, 3, size=cnt).astype(np.int8)
        miss_vals.append(v)
        for q in range(cnt):
            R.set_miss_val(j, q, int(v[q]))
    model = model_for(codes, rng)
    Ns = [0] * 5
    for snp, ti, betas, taus in model:
        R.model_add(snp, taus, ti)
        Ns[codes[ti]] += 1
    beta_e = rng.normal(size=M_E + 1) * 0.2
    beta = np.concatenate([beta_e] + [np.asarray(b) for _, _, b, _ in model])
    assert beta.size == R.model_cols()
    sigma2 = 0.7
    R.model_set_beta_sigma2(beta, sigma2)
    pp, pt = R.prior_params(), R.prior_terms()
    Ni = np.asarray(Ns, dtype=np.int32)
    L = len(model)
    lmp_add, lmp_rem = np.zeros(5), np.zeros((5, 5))
    for t in codes:
        lmp_add[t] = R.prior_log_add(Ni, L, t)
        if Ns[t] > 0:
            N2 = Ni.copy()
            N2[t] -= 1
            for u in codes:
                lmp_rem[t][u] = R.prior_log_add(N2, L - 1, u)   # sampler.cpp:61-73: "add" with one SNP of type t removed
    terms = [t for t in range(4) if cpu.term_flags(codes)[1][t]]
    tau_shared, tau_snp = np.zeros(4), np.zeros((0, len(terms)))
    if indiv:   # the scan draws one value per SNP and allowed term from the sampler's stream (sampler.cpp:99-106)
        g = ref_lib.RefRng(4242, N + 1.0)
        tau_snp = np.zeros((M_G, len(terms)))
        for j in range(M_G):
            for i in range(len(terms)):
                tau_snp[j, i] = 1.0 / (pp["alpha"] ** 2 * g.sinvchi2(pt[terms[i], 1], pt[terms[i], 2]))
    else:
        tau_shared[terms] = pt[terms, 0]
    res = R.scan()
    p_r, prt = (res if isinstance(res, tuple) else (res, np.zeros((0, 0))))
    y, E = R.y(), R.e()
    R.close()
    return dict(payload=ds["payload"], codes=codes, y=y, E=E, sigma2=sigma2, beta_e=beta_e, lmp_add=lmp_add, lmp_rem=lmp_rem,
                tau_shared=tau_shared, tau_snp=tau_snp, indiv=int(indiv), p_r=p_r, prt=prt,
                miss_vals=np.concatenate(miss_vals) if miss_vals else np.zeros(0, dtype=np.int8),
                model_snp=np.array([m[0] for m in model]), model_ti=np.array([m[1] for m in model]),
                model_beta=np.array([m[2] + [0.0] * (2 - len(m[2])) for m in model]),
                model_tau=np.array([m[3] + [0.0] * (2 - len(m[3])) for m in model]))


GOLDEN_KEYS = ("codes", "sigma2", "beta_e", "lmp_add", "lmp_rem", "tau_shared", "tau_snp", "indiv", "p_r", "prt", "miss_vals",
               "model_snp", "model_ti", "model_beta", "model_tau")


def load_case(g, tag, directory):
    """A case from the committed golden file; the data set itself is regenerated from its seed."""
    import os
    _, types, miss_rate, indiv = [c for c in CASES if c[0] == tag][0]
    os.makedirs(directory, exist_ok=True)
    ds = dataset(directory, types, miss_rate, indiv)
    c = {k: g["%s_%s" % (tag, k)] for k in GOLDEN_KEYS}
    c["codes"] = [int(v) for v in c["codes"]]
    c["sigma2"] = float(c["sigma2"])
    c["indiv"] = int(c["indiv"])
    c["payload"], c["y"] = ds["payload"], ds["y"]
    c["E"] = np.column_stack([np.ones(N), ds["E"]])
    return c


def model_terms(c):
    """The model as typed terms (an AH SNP = two terms): loci, term types, coefficients."""
    loci, tt, bg = [], [], []
    for snp, ti, b in zip(c["model_snp"], c["model_ti"], c["model_beta"]):
        code = c["codes"][int(ti)]
        if code == 4:
            loci += [int(snp), int(snp)]; tt += [0, 1]; bg += [b[0], b[1]]
        else:
            loci.append(int(snp)); tt.append(code); bg.append(b[0])
    return np.array(loci, dtype=np.int64), np.array(tt, dtype=np.int32), np.array(bg)


def oracle_scan(c, reference_offsets=True):
    off, idx, _ = None, None, None
    bed = c["payload"].copy()
    cpu.recode_minor(bed, N, M_G)
    off, idx, _ = cpu.missing_index(bed, N, M_G)
    cols0, cols = [], []
    for j in range(M_G):
        mi = idx[off[j]:off[j + 1]]
        cols0.append(cpu.decode_column_overlay(bed, N, j, 0, mi, np.zeros(mi.size, dtype=np.int8)))
        cols.append(cpu.decode_column_overlay(bed, N, j, 0, mi, c["miss_vals"][off[j]:off[j + 1]]))
    loci, tt, bg = model_terms(c)
    X = np.column_stack([c["E"]] + [cpu.typed(cols[int(j)], int(t)) for j, t in zip(loci, tt)])
    y_hat = X @ np.concatenate([c["beta_e"], bg])
    model = {int(snp): (c["codes"][int(ti)], list(b), list(t))
             for snp, ti, b, t in zip(c["model_snp"], c["model_ti"], c["model_beta"], c["model_tau"])}
    return cpu.scan_types(cols0, cols, N, c["codes"], c["y"], y_hat, model, c["sigma2"], c["lmp_add"], c["lmp_rem"],
                          tau_shared=c["tau_shared"], tau_snp=c["tau_snp"] if c["indiv"] else None,
                          reference_offsets=reference_offsets)
