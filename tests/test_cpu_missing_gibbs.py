"""CPU tests of the host side of the missing-genotype path against the UNMODIFIED
reference (oracle/_ref): the Gibbs update of the in-model SNPs' missing cells (Sampler::sample_missing,
src/sampler.cpp:264-453) and the re-imputation from the prior (DataModel::sample_missing, src/data_model.cpp:78-90).

The product's arithmetic lives in bmagwa_b200/csrc/host/missing.hpp (plain host code); tests/harness/host_harness.cpp
is compiled here with g++ and fed the same model state, the same cells and the same seed as the reference."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from bmagwa_b200 import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def harness(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("harness") / "libhost_harness.so")
    cmd = ["g++", "-O2", "-std=c++17", "-march=x86-64-v3", "-ffp-contract=off", "-fPIC", "-shared",
           "-I", os.path.join(ROOT, "bmagwa_b200", "csrc", "host"), os.path.join(ROOT, "tests", "harness", "host_harness.cpp"),
           "-o", out]
    subprocess.check_call(cmd)
    L = C.CDLL(out)
    L.harness_gibbs.restype = C.c_int
    return L


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


def _reference_state(R):
    """CSR missing index, prior and current imputed values of every SNP, from the reference."""
    off = np.zeros(R.m_g + 1, dtype=np.int64)
    idx, val, prior3 = [], [], np.zeros((R.m_g, 3))
    for j in range(R.m_g):
        i, p = R.missing(j)
        off[j + 1] = off[j] + i.size
        if i.size:
            idx.append(i)
            prior3[j] = p
            val.append(R.get_miss_val(j))
    idx = np.concatenate(idx).astype(np.int32) if idx else np.zeros(0, dtype=np.int32)
    val = np.concatenate(val).astype(np.int8) if val else np.zeros(0, dtype=np.int8)
    return off, idx, val, prior3


@pytest.mark.parametrize("seed,miss_rate,k", [(3, 0.05, 4), (4, 0.15, 7), (5, 0.30, 3)])
def test_gibbs_update_matches_the_reference(ref_lib, harness, tmp_path, seed, miss_rate, k):
    from oracle import ref
    n, m_g, m_e = 157, 60, 2
    ds = synth.write_dataset(str(tmp_path), "syn", n=n, m_g=m_g, m_e=m_e, seed=seed, miss_rate=miss_rate, e_qg=5, var_qg=20,
                             use_individual_tau2=1, do_n_iter=100, n_rao=50, n_rao_burnin=1, outbase=str(tmp_path / "chain"),
                             seeds=str(900 + seed))
    R = ref.Ref(ds["ini"])
    try:
        rng = np.random.default_rng(seed)
        loci = rng.choice(m_g, size=k, replace=False)
        # give the SNPs non-trivial imputed values first (as the prior draws of a move would)
        for j in loci:
            cnt = R.missing(int(j))[0].size
            for q in range(cnt):
                R.set_miss_val(int(j), q, int(rng.integers(0, 3)))
        for j in loci:
            R.model_add(int(j), 0.7 + 0.1 * float(rng.random()))
        cols = R.model_cols()
        assert cols == m_e + 1 + k
        beta = rng.normal(size=cols) * 0.4
        sigma2 = 0.8
        R.model_set_beta_sigma2(beta, sigma2)
        xx0, xy0 = R.model_get("xx"), R.model_get("xy")
        off, idx, val0, prior3 = _reference_state(R)
        xcols = np.asfortranarray(np.stack([R.get_column(int(j), 0, overlay=True) for j in loci], axis=1))
        y, e = R.y(), np.asfortranarray(R.e())
        assert e.shape == (n, m_e + 1)

        R.sample_missing()          # the reference: fresh stream of seed 900+seed, one uniform per cell
        xx_ref, xy_ref = R.model_get("xx"), R.model_get("xy")
        _, _, val_ref, _ = _reference_state(R)

        xx, xy, val = np.asfortranarray(xx0.copy()), xy0.copy(), val0.copy()
        q = harness.harness_gibbs(C.c_int(m_e + 1), C.c_int(k), _p(loci.astype(np.uint32), C.c_uint), _p(xx, C.c_double),
                                  _p(xy, C.c_double), _p(beta, C.c_double), C.c_double(sigma2), C.c_long(m_g), _p(off, C.c_long),
                                  _p(idx, C.c_int), _p(val, C.c_byte), _p(np.ascontiguousarray(prior3), C.c_double),
                                  _p(xcols, C.c_double), _p(y, C.c_double), _p(e, C.c_double), C.c_long(n),
                                  C.c_double(float(y @ y)), C.c_uint(900 + seed), C.c_double(n + 1.0), C.c_int(0))
        assert q > 0
        assert np.array_equal(val, val_ref), "imputed values differ from the reference's Gibbs draw"
        assert (val != val0).any(), "the update changed nothing: the test is vacuous"
        assert np.allclose(np.triu(xx), np.triu(xx_ref), rtol=1e-12, atol=1e-9)
        assert np.allclose(xy, xy_ref, rtol=1e-12, atol=1e-9)
        # the patched Gram matrix is the Gram matrix of the updated columns
        X = np.column_stack([e] + [R.get_column(int(j), 0, overlay=True) for j in loci])
        assert np.allclose(np.triu(xx), np.triu(X.T @ X), rtol=1e-12, atol=1e-9)
        assert np.allclose(xy, X.T @ y, rtol=1e-12, atol=1e-9)
    finally:
        R.close()


def test_prior_reimputation_matches_the_reference(ref_lib, harness, tmp_path):
    from oracle import ref
    n, m_g = 120, 80
    ds = synth.write_dataset(str(tmp_path), "syn", n=n, m_g=m_g, m_e=0, seed=8, miss_rate=0.1, e_qg=5, var_qg=20,
                             do_n_iter=100, n_rao=50, n_rao_burnin=1, outbase=str(tmp_path / "chain"), seeds="77")
    R = ref.Ref(ds["ini"])
    try:
        in_model = np.zeros(m_g, dtype=np.uint8)
        for j in (5, 17, 40):
            R.model_add(j, 1.0)
            in_model[j] = 1
        off, idx, val0, prior3 = _reference_state(R)
        R.sample_missing_from_prior()
        _, _, val_ref, _ = _reference_state(R)
        val = val0.copy()
        harness.harness_draw_all_from_prior(C.c_long(m_g), _p(off, C.c_long), _p(val, C.c_byte),
                                            _p(np.ascontiguousarray(prior3), C.c_double), _p(in_model, C.c_ubyte),
                                            C.c_uint(77), C.c_double(n + 1.0))
        assert np.array_equal(val, val_ref)
        for j in (5, 17, 40):   # in-model SNPs keep their values
            assert np.array_equal(val[off[j]:off[j + 1]], val0[off[j]:off[j + 1]])
        assert (val != val0).any()
    finally:
        R.close()


@pytest.mark.parametrize("types,seed,miss_rate", [("H", 21, 0.1), ("D", 22, 0.1), ("R", 23, 0.15), ("A,H,D,R", 24, 0.1),
                                                  ("AH", 25, 0.12), ("A,H,D,R,AH", 26, 0.2)])
def test_gibbs_update_with_effect_types_matches_the_reference(ref_lib, harness, tmp_path, types, seed, miss_rate):
    """The Gibbs step for models whose SNPs have effect types H, D, R and AH (two columns per SNP; sampler.cpp:318-385 and
    :393-450): imputed values, X'X and X'y against the reference, same seed."""
    from oracle import cpu, ref
    names = {"A": 0, "H": 1, "D": 2, "R": 3, "AH": 4}
    codes = sorted(names[t] for t in types.split(","))
    n, m_g, m_e = 141, 50, 1
    ds = synth.write_dataset(str(tmp_path), "syn", n=n, m_g=m_g, m_e=m_e, seed=seed, miss_rate=miss_rate, e_qg=5, var_qg=20,
                             use_individual_tau2=1, do_n_iter=100, n_rao=50, n_rao_burnin=1, types=types,
                             outbase=str(tmp_path / "chain"), seeds=str(800 + seed))
    R = ref.Ref(ds["ini"])
    try:
        rng = np.random.default_rng(seed)
        k = 6
        loci = rng.choice(m_g, size=k, replace=False)
        tis = [i % len(codes) for i in range(k)]
        for j in loci:
            for q in range(R.missing(int(j))[0].size):
                R.set_miss_val(int(j), q, int(rng.integers(0, 3)))
        for j, ti in zip(loci, tis):
            R.model_add(int(j), [0.7 + 0.1 * float(rng.random()), 0.9], ti)
        cols = R.model_cols()
        snp_type = np.array([codes[ti] for ti in tis], dtype=np.int32)
        assert cols == m_e + 1 + k + int((snp_type == 4).sum())
        beta = rng.normal(size=cols) * 0.4
        sigma2 = 0.8
        R.model_set_beta_sigma2(beta, sigma2)
        xx0, xy0 = R.model_get("xx"), R.model_get("xy")
        off, idx, val0, prior3 = _reference_state(R)
        xcols = np.asfortranarray(np.stack([R.get_column(int(j), 0, overlay=True) for j in loci], axis=1))
        y, e = R.y(), np.asfortranarray(R.e())
        R.sample_missing()
        xx_ref, xy_ref = R.model_get("xx"), R.model_get("xy")
        _, _, val_ref, _ = _reference_state(R)

        xx, xy, val = np.asfortranarray(xx0.copy()), xy0.copy(), val0.copy()
        got_cols = harness.harness_gibbs_typed(C.c_int(m_e + 1), C.c_int(k), _p(loci.astype(np.uint32), C.c_uint),
                                               _p(snp_type, C.c_int), _p(xx, C.c_double), _p(xy, C.c_double), _p(beta, C.c_double),
                                               C.c_double(sigma2), C.c_long(m_g), _p(off, C.c_long), _p(idx, C.c_int),
                                               _p(val, C.c_byte), _p(np.ascontiguousarray(prior3), C.c_double), _p(xcols, C.c_double),
                                               _p(y, C.c_double), _p(e, C.c_double), C.c_long(n), C.c_double(float(y @ y)),
                                               C.c_uint(800 + seed), C.c_double(n + 1.0))
        assert got_cols == cols
        assert np.array_equal(val, val_ref), "imputed values differ from the reference's Gibbs draw"
        assert (val != val0).any()
        assert np.allclose(np.triu(xx), np.triu(xx_ref), rtol=1e-12, atol=1e-9)
        assert np.allclose(xy, xy_ref, rtol=1e-12, atol=1e-9)
        # and the patched Gram matrix is the Gram matrix of the updated typed columns
        X = [e]
        for j, t in zip(loci, snp_type):
            a = R.get_column(int(j), 0, overlay=True)
            X += [a, cpu.typed(a, 1)] if t == 4 else [cpu.typed(a, int(t))]
        X = np.column_stack(X)
        assert np.allclose(np.triu(xx), np.triu(X.T @ X), rtol=1e-12, atol=1e-9)
    finally:
        R.close()


def test_gibbs_scratch_is_reusable_across_model_sizes(harness):
    """The sampler keeps one scratch for the whole chain: an update for a larger model after one for a smaller model must
    equal the same update on a fresh scratch, bit for bit."""
    from oracle import cpu
    n, m, k1, k, m_e = 400, 60, 5, 13, 2
    payload, f = synth.make_genotypes(n, m, seed=17, miss_rate=0.08)
    bed = payload.copy()
    cpu.recode_minor(bed, n, m)
    off, idx, prior3 = cpu.missing_index(bed, n, m)
    rs = np.random.default_rng(1)
    val = rs.integers(0, 3, size=idx.size).astype(np.int8)
    loci = rs.choice(m, size=k, replace=False).astype(np.uint32)
    cols = [cpu.decode_column_overlay(bed, n, int(j), 0, idx[off[j]:off[j + 1]], val[off[j]:off[j + 1]]) for j in loci]
    E = np.asfortranarray(np.column_stack([np.ones(n), rs.uniform(size=n)]))
    y = rs.normal(size=n)
    beta = rs.normal(size=m_e + k) * 0.2

    def gram(kk):
        X = np.column_stack([E] + cols[:kk])
        return np.asfortranarray(X.T @ X), X.T @ y

    xx1, xy1 = gram(k1)
    xx, xy = gram(k)
    xcols = np.asfortranarray(np.stack(cols, axis=1))
    ok = harness.harness_gibbs_scratch_reuse(C.c_int(m_e), C.c_int(k1), C.c_int(k), _p(loci, C.c_uint), _p(xx1, C.c_double),
                                             _p(xy1, C.c_double), _p(xx, C.c_double), _p(xy, C.c_double), _p(beta, C.c_double),
                                             C.c_double(0.9), C.c_long(m), _p(off.astype(np.int64), C.c_long),
                                             _p(idx.astype(np.int32), C.c_int), _p(val, C.c_byte),
                                             _p(np.ascontiguousarray(prior3, dtype=np.float64), C.c_double), _p(xcols, C.c_double),
                                             _p(y, C.c_double), _p(E, C.c_double), C.c_long(n), C.c_double(float(y @ y)))
    assert ok == 1
