"""CPU tests of the product's pure host classes (bmagwa_b200/csrc/host/rng.hpp, model.hpp: ChainRng, Prior, Model,
ProposalCdf) through tests/harness/host_harness.cpp, against the goldens the UNMODIFIED reference produced
(tests/golden/ref_outputs.npz) and against the reference itself when oracle/_ref is built.  The GPU chain tests exercise the
same classes end to end; these pin them without a device."""
import ctypes as C
import math
import os
import subprocess

import numpy as np
import pytest

from oracle import cpu
from tests.helpers import SMALL_INI, load_small

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def harness(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("harness") / "libhost_harness.so")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-march=x86-64-v3", "-ffp-contract=off", "-fPIC", "-shared", "-I",
                           os.path.join(ROOT, "bmagwa_b200", "csrc", "host"),
                           os.path.join(ROOT, "tests", "harness", "host_harness.cpp"), "-o", out])
    L = C.CDLL(out)
    L.harness_model_trace.restype = C.c_int
    return L


def _p(a, t=C.c_double):
    return a.ctypes.data_as(C.POINTER(t))


def test_chain_rng_reproduces_the_reference_stream(harness, golden):
    """mt19937 + uniform / normal / scaled-inverse-chi2 draws of the product's ChainRng against the reference's Rand."""
    out = np.zeros(64 * 8)
    harness.harness_rng_sequence(C.c_uint(1234), C.c_double(1001.0), C.c_int(64), _p(out))
    assert np.allclose(out, golden["rng_seed1234_nu1001"], rtol=1e-14, atol=0)


# prior of SMALL_INI (tests/helpers.py): e_qg 2, var_qg 2, R2mode_sigma2 0.2, nu_sigma2 1, nu_tau2_A 3, s2_tau2_A 0.02
def _small_prior_args(y, n=30, m_g=7):
    var_y = float(np.var(y, ddof=1))
    nu = 1.0
    s2_sigma2 = var_y * (1 - 0.2) * (nu + 2) / nu   # sampler.hpp:173-177
    return dict(n=n, m_g=m_g, m_e=1, yy=float(y @ y), e_qg=2.0, var_qg=2.0, nu_sigma2=nu, s2_sigma2=s2_sigma2, nu_tau2=3.0,
                s2_tau2=0.02)


def test_prior_matches_reference(harness, golden, golden_data_dir):
    bed, y = load_small(golden_data_dir)
    a = _small_prior_args(y)
    out5, add, rem, model = np.zeros(5), np.zeros(7), np.zeros(7), np.zeros(7)
    harness.harness_prior(C.c_long(a["n"]), C.c_long(a["m_g"]), C.c_int(1), C.c_double(a["yy"]), C.c_double(a["e_qg"]),
                          C.c_double(a["var_qg"]), C.c_double(a["nu_sigma2"]), C.c_double(a["s2_sigma2"]), C.c_double(a["nu_tau2"]),
                          C.c_double(a["s2_tau2"]), C.c_double(1.0), C.c_int(0), C.c_int(7), _p(out5), _p(add), _p(rem), _p(model))
    g_a, g_b, n_plus_nu, nus2_plus_yy, alpha, s2_sigma2, nu_tau2, s2_tau2, shared, e_g = golden["small_prior"]
    assert out5[0] == n_plus_nu and out5[1] == pytest.approx(nus2_plus_yy, rel=1e-13)
    assert out5[2] == alpha and out5[3] == pytest.approx(shared, rel=1e-14) and out5[4] == pytest.approx(e_g, rel=1e-13)
    # log ratios are differences of the full log model prior (prior.hpp:144-183), as the oracle's Prior
    P = cpu.Prior.make(7, 2, 2)
    for L in range(6):
        assert add[L] == pytest.approx(P.log_add([L, 0, 0, 0, 0], L), rel=1e-12, abs=1e-12)
        assert rem[L] == pytest.approx(P.log_rem([L + 1, 0, 0, 0, 0], L + 1), rel=1e-12, abs=1e-12)
        assert add[L] == pytest.approx(model[L + 1] - model[L], rel=1e-10, abs=1e-10)
        assert rem[L] == pytest.approx(model[L] - model[L + 1], rel=1e-10, abs=1e-10)


def test_prior_log_ratios_against_live_reference(harness, ref_lib, tmp_path, golden_data_dir):
    ini = tmp_path / "c.ini"
    ini.write_text(SMALL_INI.format(d=golden_data_dir, out=str(tmp_path), recode=1, do_n_iter=10, n_rao=10, n_rao_burnin=1,
                                    delay_rejection=7, thin=10, seed=1, indiv=0))
    R = ref_lib.Ref(str(ini))
    bed, y = load_small(golden_data_dir)
    a = _small_prior_args(y)
    out5, add, rem, model = np.zeros(5), np.zeros(7), np.zeros(7), np.zeros(7)
    harness.harness_prior(C.c_long(a["n"]), C.c_long(a["m_g"]), C.c_int(1), C.c_double(a["yy"]), C.c_double(a["e_qg"]),
                          C.c_double(a["var_qg"]), C.c_double(a["nu_sigma2"]), C.c_double(a["s2_sigma2"]), C.c_double(a["nu_tau2"]),
                          C.c_double(a["s2_tau2"]), C.c_double(1.0), C.c_int(0), C.c_int(7), _p(out5), _p(add), _p(rem), _p(model))
    for L in range(6):
        Ns = np.array([L, 0, 0, 0, 0], dtype=np.int32)
        assert add[L] == pytest.approx(R.prior_log_add(Ns, L, 0), rel=1e-12, abs=1e-12)
        Ns1 = np.array([L + 1, 0, 0, 0, 0], dtype=np.int32)
        assert rem[L] == pytest.approx(R.prior_log_rem(Ns1, L + 1, 0), rel=1e-12, abs=1e-12)
        # log_model keeps the L-dependent terms only (what the sampler writes to _log_prior.dat)
        assert model[L] == pytest.approx(R.prior_log_model(Ns), rel=1e-12, abs=1e-10)
    R.close()


def test_model_add_remove_trace_matches_reference(harness, golden, golden_data_dir):
    """Model::add_term / remove_term (rank-1 Cholesky append, dchex-style delete, O(1) likelihood updates) against the
    reference's own trace of the same operations, and its final Gram matrix, factor and v."""
    bed, y = load_small(golden_data_dir)
    n, m_g = 30, 7
    a = _small_prior_args(y)
    G = np.asfortranarray(np.stack([cpu.decode_column(bed, n, j, 0) for j in range(m_g)], axis=1))
    E = np.asfortranarray(np.ones((n, 1)))
    ops = np.ascontiguousarray(golden["ll_ops"], dtype=np.float64)
    n_ops = ops.shape[0]
    trace = np.zeros(n_ops + 1)
    cols_max = 1 + m_g
    xx, l = np.zeros(cols_max * cols_max), np.zeros(cols_max * cols_max)
    xy, v, full = np.zeros(cols_max), np.zeros(cols_max), np.zeros(1)
    loci = np.zeros(m_g, dtype=np.uint32)
    cols = harness.harness_model_trace(C.c_long(n), C.c_long(m_g), C.c_int(1), _p(G), _p(E), _p(y), C.c_double(a["yy"]),
                                       C.c_double(a["e_qg"]), C.c_double(a["var_qg"]), C.c_double(a["nu_sigma2"]),
                                       C.c_double(a["s2_sigma2"]), C.c_double(a["nu_tau2"]), C.c_double(a["s2_tau2"]), C.c_int(n_ops),
                                       _p(ops.reshape(-1)), _p(trace), _p(xx), _p(l), _p(xy), _p(v), _p(full), _p(loci, C.c_uint))
    assert np.allclose(trace, golden["ll_trace"], rtol=0, atol=1e-10)
    k = cols - 1
    assert list(loci[:k]) == list(golden["ll_final_loci"])
    assert np.allclose(xx[:cols * cols].reshape(cols, cols).T, golden["ll_final_xx"], rtol=0, atol=1e-12)
    assert np.allclose(l[:cols * cols].reshape(cols, cols).T, golden["ll_final_l"], rtol=0, atol=1e-10)
    assert np.allclose(xy[:cols], golden["ll_final_xy"], rtol=0, atol=1e-12)
    assert np.allclose(v[:cols], golden["ll_final_v"], rtol=0, atol=1e-10)
    assert full[0] == pytest.approx(golden["ll_final_full"][0], abs=1e-10)
    assert trace[-1] == pytest.approx(full[0], abs=1e-10)   # incremental = full recomputation


@pytest.mark.parametrize("m,block", [(1, 256), (2, 1), (5, 2), (10, 4), (31, 8), (257, 16), (257, 256)])
def test_proposal_cdf_matches_reference_tree_draws(harness, golden, m, block):
    """ProposalCdf (in-order layout + per-block partial CDFs + Fenwick tree) draws the items the reference's
    DiscreteDistribution tree draws from the same uniforms, through zero / unzero updates."""
    w = np.ascontiguousarray(golden["dd_w_%d" % m])
    rec = np.ascontiguousarray(golden["dd_rec_%d" % m], dtype=np.float64)
    order = np.ascontiguousarray(cpu.inorder_permutation(m), dtype=np.int32)
    got = np.zeros(rec.shape[0], dtype=np.int64)
    total = np.zeros(rec.shape[0])
    harness.harness_proposal_cdf(C.c_long(m), _p(order, C.c_int), _p(w), C.c_int(block), C.c_int(rec.shape[0]), _p(rec.reshape(-1)),
                                 _p(got, C.c_long), _p(total))
    for i, (kind, item, tot, u) in enumerate(rec):
        if kind == 2:
            assert got[i] == int(item), "draw %d" % i
        assert total[i] == pytest.approx(tot, rel=1e-12)


@pytest.mark.parametrize("const_loci,ms,seed", [(0, 2, 1), (2, 3, 2), (3, 5, 3), (1, 7, 4), (0, 1, 5)])
def test_exhaustive_modelset_equals_brute_force(harness, const_loci, ms, seed):
    """Delayed rejection enumerates the 2^ms sub-models of a move's SNPs by adjacent Givens swaps of the Cholesky factor
    (sampler.cpp:882-980, model.hpp:585-844): every log probability must equal the one of a model built from scratch."""
    from bmagwa_b200 import synth
    n, m_g = 120, 40
    payload, f = synth.make_genotypes(n, m_g, seed=seed)
    bed = payload.copy()
    cpu.recode_minor(bed, n, m_g)
    rs = np.random.default_rng(seed)
    G = np.asfortranarray(np.stack([cpu.decode_column(bed, n, j, 0) for j in range(m_g)], axis=1))
    E = np.asfortranarray(np.column_stack([np.ones(n), rs.uniform(size=n)]))
    y = rs.normal(size=n) + 0.5 * G[:, 3]
    snps = rs.choice(m_g, size=const_loci + ms, replace=False).astype(np.uint32)
    taus = 0.5 + rs.random(size=const_loci + ms) * 5
    P, B = np.zeros(1 << ms), np.zeros(1 << ms)
    var_y = float(np.var(y, ddof=1))
    harness.harness_exhaustive(C.c_long(n), C.c_long(m_g), C.c_int(2), _p(G), _p(E), _p(y), C.c_double(float(y @ y)), C.c_double(5.0),
                               C.c_double(20.0), C.c_double(1.0), C.c_double(var_y * 0.8 * 3.0), C.c_double(5.0), C.c_double(0.05),
                               C.c_int(const_loci), C.c_int(ms), _p(snps, C.c_uint), _p(taus), _p(P), _p(B))
    assert np.isfinite(P).all() and np.isfinite(B).all()
    assert np.allclose(P, B, rtol=0, atol=1e-9)
    assert np.abs(B).max() > 1e-3


@pytest.mark.parametrize("with_types", [False, True])
@pytest.mark.parametrize("n_inds,const_loci,m_g,seed", [(2, 0, 50, 1), (3, 4, 50, 2), (6, 1, 30, 3), (8, 10, 200, 4), (4, 0, 4, 5)])
def test_dr_proposal_probabilities_match_reference_function(harness, ref_lib, n_inds, const_loci, m_g, seed, with_types):
    """compute_proposal_probs_for_exh_modelset: the product's restatement (logs of the weights taken once, normalising totals
    multiplied up and logged once per sub-model) against the reference's own function (sampler.cpp:982-1049)."""
    rs = np.random.default_rng(seed)
    order = rs.permutation(n_inds).astype(np.uint8)
    q_add = rs.uniform(0.01, 0.5, size=n_inds)
    q_rem = rs.uniform(0.5, 1.0, size=n_inds)
    z_add = q_add.sum() + rs.uniform(1.0, 5.0)
    z_rem = rs.uniform(0.5, 3.0) * (1 if const_loci else 0) + 1e-300 * 0
    if const_loci == 0:
        z_rem = 0.0
    start = rs.normal(size=1 << n_inds)
    ours, ref = start.copy(), start.copy()
    lqt = np.log(rs.uniform(0.1, 1.0, size=n_inds)) if with_types else None   # type proposals of the additions (several types)
    harness.harness_dr_proposal_probs(C.c_int(n_inds), _p(order, C.c_ubyte), _p(q_add), _p(q_rem), C.c_double(z_add), C.c_double(z_rem),
                                      C.c_long(const_loci), C.c_long(m_g), _p(lqt) if with_types else None, _p(ours))
    ref_lib.lib().refd_dr_proposal_probs(C.c_int(n_inds), _p(order, C.c_ubyte), _p(q_add), _p(q_rem), C.c_double(z_add),
                                         C.c_double(z_rem), C.c_long(const_loci), C.c_long(m_g), _p(lqt) if with_types else None,
                                         _p(ref))
    assert np.isfinite(ref).all()
    assert np.allclose(ours, ref, rtol=1e-12, atol=1e-11)


@pytest.mark.parametrize("indiv,k,seed", [(1, 4, 11), (0, 6, 12), (1, 0, 13)])
def test_model_level_gibbs_updates_match_reference(harness, ref_lib, tmp_path, indiv, k, seed):
    """sample_beta_sigma2, sample_alpha_and_tau2, compute_log_likelihood and compute_pve of the product (Gram-matrix
    formulation, no n-vectors) against the reference's (n x k design matrix), same seed: the draws must coincide."""
    from bmagwa_b200 import synth
    n, m_g, m_e = 203, 120, 2
    ds = synth.write_dataset(str(tmp_path), "syn", n=n, m_g=m_g, m_e=m_e, seed=seed, e_qg=5, var_qg=20, use_individual_tau2=indiv,
                             do_n_iter=100, n_rao=50, n_rao_burnin=1, outbase=str(tmp_path / "chain"), seeds=str(700 + seed))
    R = ref_lib.Ref(ds["ini"])
    try:
        rs = np.random.default_rng(seed)
        snps = rs.choice(m_g, size=k, replace=False).astype(np.uint32) if k else np.zeros(0, dtype=np.uint32)
        taus = 0.5 + rs.random(size=max(k, 1)) * 4
        for j, t in zip(snps, taus):
            R.model_add(int(j), float(t))
        cols = R.model_cols()
        pp, pt = R.prior_params(), R.prior_terms()
        y, E = R.y(), np.asfortranarray(R.e())
        G = np.asfortranarray(np.stack([R.get_column(j, 0) for j in range(m_g)], axis=1))
        R.model_sample_beta_sigma2()
        beta_ref = R.model_get("beta")[:cols].copy()
        sigma2_ref = R.model_get("scalars")["sigma2"]
        R.sample_alpha_and_tau2()
        tau_ref = R.model_get("inv_tau2_alpha2")[:cols].copy()
        alpha_ref = R.prior_params()["alpha"]
        R.model_compute_loglik()
        ll_ref = R.model_loglik()
        pves_ref, _ = R.model_compute_pve()

        beta, tau, out3, pves = np.zeros(cols), np.zeros(cols), np.zeros(3), np.zeros(3)
        harness.harness_model_gibbs(C.c_long(n), C.c_long(m_g), C.c_int(m_e + 1), _p(G), _p(E), _p(y), C.c_double(float(y @ y)),
                                    C.c_double(5.0), C.c_double(20.0), C.c_double(1.0), C.c_double(pp["s2_sigma2"]),
                                    C.c_double(pt[0, 1]), C.c_double(pt[0, 2]), C.c_double(1.0), C.c_int(indiv), C.c_double(0.001),
                                    C.c_int(k), _p(snps, C.c_uint), _p(taus), C.c_uint(700 + seed), _p(beta), _p(tau), _p(out3), _p(pves))
        assert np.allclose(beta, beta_ref, rtol=1e-9, atol=1e-11)
        assert out3[0] == pytest.approx(sigma2_ref, rel=1e-11)
        assert out3[1] == pytest.approx(alpha_ref, rel=1e-9)
        assert np.allclose(tau[m_e + 1:], tau_ref[m_e + 1:], rtol=1e-9)
        assert out3[2] == pytest.approx(ll_ref, rel=1e-10)
        assert np.allclose(pves, pves_ref, rtol=1e-8, atol=1e-12)
    finally:
        R.close()


# ------------------------------------------------------------------------------- several effect types
NAMES = {"A": 0, "H": 1, "D": 2, "R": 3, "AH": 4}


@pytest.mark.parametrize("types,seed", [("A,H", 31), ("A,H,D,R", 32), ("AH", 33), ("A,H,D,R,AH", 34), ("D", 35)])
def test_typed_prior_matches_reference(harness, ref_lib, tmp_path, types, seed):
    """Prior::log_change_on_add / _rem / _swi / log_model with per-type counts Ns, and the shared per-term precisions, against
    the reference's Prior (prior.hpp:60-183) for several model.types settings."""
    from bmagwa_b200 import synth
    n, m_g, m_e = 101, 40, 1
    ds = synth.write_dataset(str(tmp_path), "syn", n=n, m_g=m_g, m_e=m_e, seed=seed, e_qg=5, var_qg=20, use_individual_tau2=0,
                             do_n_iter=100, n_rao=50, n_rao_burnin=1, types=types, outbase=str(tmp_path / "chain"), seeds="1")
    R = ref_lib.Ref(ds["ini"])
    try:
        codes = sorted(NAMES[t] for t in types.split(","))
        rs = np.random.default_rng(seed)
        pp, pt = R.prior_params(), R.prior_terms()
        queries, swi_rem = [], []
        for _ in range(25):
            Ns = np.zeros(5, dtype=np.int32)
            for t in codes:
                Ns[t] = rs.integers(0, 5)
            queries.append(list(Ns) + [int(Ns.sum()), int(rs.choice(codes))])
            swi_rem.append(int(rs.choice(codes)))
        q = np.ascontiguousarray(queries, dtype=np.int32)
        sr = np.ascontiguousarray(swi_rem, dtype=np.int32)
        add, rem, model, swi = (np.zeros(len(queries)) for _ in range(4))
        shared4 = np.zeros(4)
        y = R.y()
        harness.harness_prior_typed(C.c_long(n), C.c_long(m_g), C.c_int(m_e + 1), C.c_double(float(y @ y)), C.c_double(5.0),
                                    C.c_double(20.0), C.c_double(pp["s2_sigma2"]), C.c_int(len(codes)),
                                    _p(np.asarray(codes, dtype=np.int32), C.c_int), C.c_int(len(queries)), _p(q.reshape(-1), C.c_int),
                                    _p(sr, C.c_int), _p(add), _p(rem), _p(model), _p(swi), _p(shared4))
        for i, row in enumerate(queries):
            Ns, L, t = np.asarray(row[:5], dtype=np.int32), row[5], row[6]
            assert add[i] == pytest.approx(R.prior_log_add(Ns, L, t), rel=1e-13, abs=1e-13)
            if Ns[t] > 0:
                assert rem[i] == pytest.approx(R.prior_log_rem(Ns, L, t), rel=1e-13, abs=1e-13)
            assert model[i] == pytest.approx(R.prior_log_model(Ns), rel=1e-13, abs=1e-12)
            if Ns[swi_rem[i]] > 0:
                assert swi[i] == pytest.approx(R.prior_log_swi(Ns, t, swi_rem[i]), rel=1e-13, abs=1e-13)
        for t in range(4):
            if np.isnan(pt[t, 0]):
                assert shared4[t] == -1.0
            else:
                assert shared4[t] == pytest.approx(pt[t, 0], rel=1e-14)
    finally:
        R.close()


@pytest.mark.parametrize("types,indiv,seed", [("A,H,D,R", 1, 41), ("A,H,D,R", 0, 42), ("AH", 0, 43), ("A,H,D,R,AH", 1, 44),
                                              ("H,R", 0, 45)])
def test_typed_model_gibbs_updates_match_reference(harness, ref_lib, tmp_path, types, indiv, seed):
    """A model whose SNPs have effect types (AH = two columns): sample_beta_sigma2 and the typed tau2 / alpha update
    (prior.cpp:71-141 with x_types; one tau2 per term type in the shared mode) against the reference, same seed."""
    from bmagwa_b200 import synth
    n, m_g, m_e = 151, 60, 1
    ds = synth.write_dataset(str(tmp_path), "syn", n=n, m_g=m_g, m_e=m_e, seed=seed, e_qg=5, var_qg=20, use_individual_tau2=indiv,
                             do_n_iter=100, n_rao=50, n_rao_burnin=1, types=types, outbase=str(tmp_path / "chain"),
                             seeds=str(600 + seed))
    R = ref_lib.Ref(ds["ini"])
    try:
        codes = sorted(NAMES[t] for t in types.split(","))
        rs = np.random.default_rng(seed)
        k = 6
        snps = rs.choice(m_g, size=k, replace=False).astype(np.uint32)
        tis = [i % len(codes) for i in range(k)]
        snp_type = np.array([codes[ti] for ti in tis], dtype=np.int32)
        taus2 = 0.5 + rs.random(size=(k, 2)) * 3
        for j, ti, t2 in zip(snps, tis, taus2):
            R.model_add(int(j), list(t2), ti)
        cols = R.model_cols()
        pp = R.prior_params()
        y, E = R.y(), np.asfortranarray(R.e())
        G = np.asfortranarray(np.stack([R.get_column(j, 0) for j in range(m_g)], axis=1))
        R.model_sample_beta_sigma2()
        beta_ref = R.model_get("beta")[:cols].copy()
        sigma2_ref = R.model_get("scalars")["sigma2"]
        R.sample_alpha_and_tau2()
        tau_ref = R.model_get("inv_tau2_alpha2")[:cols].copy()
        alpha_ref = R.prior_params()["alpha"]
        R.model_compute_loglik()
        ll_ref = R.model_loglik()

        beta, tau, out3 = np.zeros(cols), np.zeros(cols), np.zeros(3)
        got = harness.harness_model_gibbs_typed(C.c_long(n), C.c_long(m_g), C.c_int(m_e + 1), _p(G), _p(E), _p(y),
                                                C.c_double(float(y @ y)), C.c_double(5.0), C.c_double(20.0),
                                                C.c_double(pp["s2_sigma2"]), C.c_int(indiv), C.c_int(len(codes)),
                                                _p(np.asarray(codes, dtype=np.int32), C.c_int), C.c_int(k), _p(snps, C.c_uint),
                                                _p(snp_type, C.c_int), _p(np.ascontiguousarray(taus2).reshape(-1)),
                                                C.c_uint(600 + seed), _p(beta), _p(tau), _p(out3))
        assert got == cols
        assert np.allclose(beta, beta_ref, rtol=1e-9, atol=1e-11)
        assert out3[0] == pytest.approx(sigma2_ref, rel=1e-11)
        assert out3[1] == pytest.approx(alpha_ref, rel=1e-9)
        assert np.allclose(tau[m_e + 1:], tau_ref[m_e + 1:], rtol=1e-9)
        assert out3[2] == pytest.approx(ll_ref, rel=1e-10)
    finally:
        R.close()


@pytest.mark.parametrize("types,seed", [("A,H,D,R,AH", 51), ("AH", 52), ("A,AH", 53)])
def test_typed_model_add_remove_trace_matches_reference(harness, ref_lib, tmp_path, types, seed):
    """Adding and removing SNPs of several effect types (AH = two columns, removed larger column first) through TypedTerms +
    Model: the log-likelihood after every operation and the final X'X against the reference's Model (model.hpp:239-312)."""
    from bmagwa_b200 import synth
    n, m_g, m_e = 131, 40, 1
    ds = synth.write_dataset(str(tmp_path), "syn", n=n, m_g=m_g, m_e=m_e, seed=seed, e_qg=5, var_qg=20, use_individual_tau2=1,
                             do_n_iter=100, n_rao=50, n_rao_burnin=1, types=types, outbase=str(tmp_path / "chain"), seeds="3")
    R = ref_lib.Ref(ds["ini"])
    try:
        codes = sorted(NAMES[t] for t in types.split(","))
        rs = np.random.default_rng(seed)
        in_model, ops, trace_ref = [], [], []
        free = list(rs.permutation(m_g))
        for step in range(24):
            if in_model and (rs.random() < 0.4 or len(in_model) > 8):
                mi = int(rs.integers(0, len(in_model)))
                R.model_remove(mi)
                in_model.pop(mi)
                ops.append([1, mi, 0, 0, 0])
            else:
                snp, ti = int(free.pop()), int(rs.integers(0, len(codes)))
                t1, t2 = 0.5 + 3 * rs.random(), 0.5 + 3 * rs.random()
                R.model_add(snp, [t1, t2], ti)
                in_model.append(snp)
                ops.append([0, snp, codes[ti], t1, t2])
            trace_ref.append(R.model_loglik())
        cols = R.model_cols()
        xx_ref = R.model_get("xx")
        pp = R.prior_params()
        y, E = R.y(), np.asfortranarray(R.e())
        G = np.asfortranarray(np.stack([R.get_column(j, 0) for j in range(m_g)], axis=1))
        ops_a = np.ascontiguousarray(ops, dtype=np.float64)
        trace = np.zeros(len(ops))
        xx = np.zeros(64 * 64)
        Ns = np.zeros(5, dtype=np.int32)
        got = harness.harness_typed_model_trace(C.c_long(n), C.c_long(m_g), C.c_int(m_e + 1), _p(G), _p(E), _p(y), C.c_double(float(y @ y)),
                                                C.c_double(pp["s2_sigma2"]), C.c_int(len(codes)),
                                                _p(np.asarray(codes, dtype=np.int32), C.c_int), C.c_int(len(ops)), _p(ops_a.reshape(-1)),
                                                _p(trace), _p(xx), _p(Ns, C.c_int))
        assert got == cols
        assert np.allclose(trace, trace_ref, rtol=0, atol=1e-9)
        assert np.allclose(xx[:cols * cols].reshape(cols, cols).T, xx_ref, rtol=0, atol=1e-10)
        assert Ns.sum() == len(in_model)
    finally:
        R.close()


@pytest.mark.parametrize("types,const_loci,ms,seed", [("A,H,D,R", 2, 4, 61), ("AH", 1, 3, 62), ("A,AH", 2, 5, 63),
                                                       ("A,H,D,R,AH", 3, 6, 64), ("A,H,D,R,AH", 0, 7, 65)])
def test_typed_exhaustive_modelset_equals_brute_force(harness, types, const_loci, ms, seed):
    """TypedExhModel walks the 2^ms sub-models of SNPs with effect types (an AH SNP moves as a pair of columns: one, two or
    four adjacent Givens swaps per step, model.hpp:706-829): every log probability -- likelihood plus the per-type model
    prior -- must equal the one of a model built from scratch."""
    from bmagwa_b200 import synth
    n, m_g = 150, 40
    payload, f = synth.make_genotypes(n, m_g, seed=seed)
    bed = payload.copy()
    cpu.recode_minor(bed, n, m_g)
    rs = np.random.default_rng(seed)
    codes = sorted(NAMES[t] for t in types.split(","))
    G = np.asfortranarray(np.stack([cpu.decode_column(bed, n, j, 0) for j in range(m_g)], axis=1))
    E = np.asfortranarray(np.column_stack([np.ones(n), rs.uniform(size=n)]))
    y = rs.normal(size=n) + 0.5 * G[:, 3]
    k = const_loci + ms
    snps = rs.choice(m_g, size=k, replace=False).astype(np.uint32)
    snp_type = np.array([codes[int(rs.integers(0, len(codes)))] for _ in range(k)], dtype=np.int32)
    if 4 in codes:
        snp_type[const_loci] = 4          # at least one pair among the moving SNPs ...
        if ms > 1:
            snp_type[const_loci + 1] = 4  # ... and two pairs next to each other
    taus2 = 0.5 + rs.random(size=(k, 2)) * 5
    P, B = np.zeros(1 << ms), np.zeros(1 << ms)
    var_y = float(np.var(y, ddof=1))
    harness.harness_exhaustive_typed(C.c_long(n), C.c_long(m_g), C.c_int(2), _p(G), _p(E), _p(y), C.c_double(float(y @ y)),
                                     C.c_double(var_y * 0.8 * 3.0), C.c_int(len(codes)), _p(np.asarray(codes, dtype=np.int32), C.c_int),
                                     C.c_int(const_loci), C.c_int(ms), _p(snps, C.c_uint), _p(snp_type, C.c_int),
                                     _p(np.ascontiguousarray(taus2).reshape(-1)), _p(P), _p(B))
    assert np.isfinite(P).all() and np.isfinite(B).all()
    assert np.allclose(P, B, rtol=0, atol=1e-9)
    assert np.abs(B).max() > 1e-3


def test_memo_pair_table_through_growth_and_restarts(harness):
    """gramcache.hpp: the open-addressing table of x_j'x_l products never returns a wrong value -- while it grows (no
    entry lost) and once it is full and starts over (entries gone, not garbled).  A chain meets the second case after some
    10^5 iterations on a small SNP panel; the sampler must not rely on an entry surviving the filing of a move's results
    (it does not: bmagwa_b200/csrc/host/sampler.cpp, finish_gram)."""
    harness.harness_gramcache.restype = C.c_long
    out = np.zeros(3)
    # growing only: 200,000 pairs fit below the default limit, nothing may be lost
    assert harness.harness_gramcache(C.c_long(0), C.c_long(200000), C.c_long(5000), _p(out)) == 0
    assert out[0] == 0 and out[1] > 190000 and out[2] == 200000
    # 4,096 slots: the table starts over every ~2,048 distinct pairs
    assert harness.harness_gramcache(C.c_long(4096), C.c_long(100000), C.c_long(5000), _p(out)) == 0
    assert out[0] >= 40 and out[1] <= 2048 and 0 < out[2] <= 4096
    # degenerate limits are rounded to a usable table
    assert harness.harness_gramcache(C.c_long(1), C.c_long(1000), C.c_long(50), _p(out)) == 0 and out[0] > 0
