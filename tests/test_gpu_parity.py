"""GPU parity tests (run on the B200 box: pytest -m gpu).  Every check calls the CUDA path through
the C ABI (bmagwa_b200.api over libbmagwa_b200.so) and compares with the CPU oracle
(oracle/oracle.c) on the same seeded inputs, or with the committed reference goldens.

Bars: bit-exact for genotype decode, counts, missing index, recode flags and the moment cache;
x_j'r to 1e-12 relative to ||x_j||*||r|| and p_r to 1e-9 absolute (north star: 1e-9; H7 explains
why the dot product is compared against the norm product, not its own magnitude)."""
import math
import os

import numpy as np
import pytest

from bmagwa_b200 import synth
from oracle import cpu
from tests.helpers import load_plink, load_small

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def api():
    from bmagwa_b200 import api as _api
    return _api


def make_data(n, m, seed, miss_rate=0.0, m_e=2):
    payload, f = synth.make_genotypes(n, m, seed=seed, miss_rate=miss_rate)
    y, causal, beta = synth.make_phenotype(payload, f, n, m, seed=seed, n_causal=min(20, m))
    rs = np.random.default_rng(seed + 2)
    E = rs.uniform(size=(n, m_e))
    return payload, y, E


def oracle_store(payload, n, m, recode):
    bed = payload.copy()
    sw = cpu.recode_minor(bed, n, m) if recode else np.zeros(m, dtype=np.uint8)
    return bed, sw


# ------------------------------------------------------------------------------- store (a1-a4)
@pytest.mark.parametrize("n,m,miss,recode", [(30, 7, 0.0, True), (203, 300, 0.01, True), (1000, 257, 0.0, False),
                                             (5000, 64, 0.002, True), (16, 5, 0.2, True), (1, 3, 0.0, False),
                                             (33, 40, 0.5, True)])
def test_store_decode_counts_moments_missing_bit_exact(api, n, m, miss, recode):
    payload, y, E = make_data(n, m, seed=n + m, miss_rate=miss)
    bed, sw = oracle_store(payload, n, m, recode)
    st = api.GenotypeStore(payload, n, m, recode_to_minor=recode)
    st.set_phenotype(y, E)
    n1, n2, nm, swapped = st.counts()
    assert np.array_equal(swapped, sw)
    G = cpu.decode_matrix(bed, n, m, 0)
    assert np.array_equal(n1, (G == 1).sum(axis=0)) and np.array_equal(n2, (G == 2).sum(axis=0))
    assert np.array_equal(nm, (G == -1).sum(axis=0))
    for j in range(0, m, max(1, m // 23)):
        for t in range(4):
            assert np.array_equal(st.get_column(j, t), cpu.decode_column(bed, n, j, t)), (j, t)
    assert np.array_equal(st.moments(), cpu.moments(bed, n, m))
    off, idx, prior = st.missing()
    off2, idx2, prior2 = cpu.missing_index(bed, n, m)
    assert np.array_equal(off, off2) and np.array_equal(idx, idx2)
    has = (off[1:] - off[:-1]) > 0
    assert np.array_equal(prior[has], prior2[has])
    if n > 1:
        mean, var = cpu.g_var_and_mean(bed, n, m)
        s = st.summaries()
        assert s["mean_x"] == mean and (s["var_x"] == var or (math.isnan(var) and math.isnan(s["var_x"])))
        assert s["var_y"] == pytest.approx(cpu.var(y), rel=1e-14)
    st.close()


def test_store_reference_fixtures_bit_exact(api, ref_tests, golden, golden_data_dir):
    """The reference's own decode goldens (src/tests/data_tests.hpp:40-47,130-140,171-190)."""
    raw = cpu.read_bed(os.path.join(golden_data_dir, "plinktest.bed"), 5, 10)
    st = api.GenotypeStore(raw, 5, 10, recode_to_minor=False)
    G = np.array(ref_tests["plinktest"]["genotypes_snp_major"], dtype=np.float64)
    for j in range(10):
        assert np.array_equal(st.get_column(j, 0), G[j])
        for t in range(1, 4):
            assert np.array_equal(st.get_column(j, t), golden["plink_cols_raw"][t][:, j])
    off, idx, prior = st.missing()
    assert list(idx[off[0]:off[1]]) == [3] and list(prior[0]) == [1, 2, 4]
    assert list(idx[off[3]:off[4]]) == [1] and list(prior[3]) == [2, 3, 4]
    assert np.array_equal(st.moments(), golden["plink_moments"])
    # overlay (data_model.cpp:30-72): SNP0/ind3 := 2, SNP3/ind1 := 1
    _, y = load_plink(golden_data_dir)
    st.set_phenotype(y)
    ch = api.Chain(st)
    ch.set_missing(0, [2])
    ch.set_missing(3, [1])
    for t in range(4):
        for j in range(10):
            assert np.array_equal(ch.get_column(j, t), golden["plink_cols_overlay"][t][:, j])
    ch.close()
    st.close()
    raw = cpu.read_bed(os.path.join(golden_data_dir, "small_modelspace.bed"), 30, 7)
    st = api.GenotypeStore(raw, 30, 7, recode_to_minor=True)
    for j in range(7):
        assert np.array_equal(st.get_column(j, 0), golden["small_cols_A"][:, j])
    assert np.array_equal(st.moments(), golden["small_moments"])
    st.close()


def test_store_streamed_from_bed_file_equals_store_from_payload(api, golden_data_dir, tmp_path):
    """bmg_store_create_from_bed (file -> pinned staging -> device) builds the same store as the in-memory
    payload path, for the whole file, for an SNP shard and for a file larger than one staging block; the
    header checks carry the reference's messages (data.cpp:250-262)."""
    from bmagwa_b200 import synth
    path = os.path.join(golden_data_dir, "plinktest.bed")
    raw = cpu.read_bed(path, 5, 10)
    a = api.GenotypeStore(raw, 5, 10, recode_to_minor=True)
    b = api.GenotypeStore(None, 5, 10, recode_to_minor=True, bed_path=path)
    for j in range(10):
        assert np.array_equal(a.get_column(j, 0), b.get_column(j, 0))
    assert np.array_equal(a.moments(), b.moments())
    for x, y in zip(a.missing(), b.missing()):
        assert np.array_equal(x, y)
    a.close(); b.close()
    n, m = 4001, 20000   # 1001 bytes per SNP -> 20 MB payload = many staging blocks, ragged last byte
    payload, _ = synth.make_genotypes(n, m, seed=9, miss_rate=0.01)
    big = tmp_path / "big.bed"
    with open(big, "wb") as fh:
        fh.write(bytes([0x6C, 0x1B, 0x01]))
        fh.write(payload.tobytes())
    a = api.GenotypeStore(payload.reshape(m, -1)[5000:17000].copy(), n, m, recode_to_minor=True, snp_lo=5000, snp_hi=17000)
    b = api.GenotypeStore(None, n, m, recode_to_minor=True, snp_lo=5000, snp_hi=17000, bed_path=str(big))
    assert all(np.array_equal(x, y) for x, y in zip(a.counts(), b.counts()))
    assert np.array_equal(a.moments(), b.moments())
    for j in (5000, 11111, 16999):
        assert np.array_equal(a.get_column(j, 0), b.get_column(j, 0))
    for x, y in zip(a.missing(), b.missing()):
        assert np.array_equal(x, y)
    a.close(); b.close()
    bad = tmp_path / "bad.bed"
    bad.write_bytes(bytes([0x6C, 0x1B, 0x00]) + payload.tobytes()[:100])
    with pytest.raises(RuntimeError, match="BED file not in snp-major format"):
        api.GenotypeStore(None, n, m, bed_path=str(bad))
    with pytest.raises(RuntimeError, match="BED file could not be opened"):
        api.GenotypeStore(None, n, m, bed_path=str(tmp_path / "absent.bed"))
    short = tmp_path / "short.bed"
    short.write_bytes(bytes([0x6C, 0x1B, 0x01]) + payload.tobytes()[:1000])
    with pytest.raises(RuntimeError, match="Reading the BED file failed"):
        api.GenotypeStore(None, n, m, bed_path=str(short))


# ------------------------------------------------------------------- residual + scan (a5, a9)
def random_state(n, m, m_e, k, seed):
    rs = np.random.default_rng(seed)
    loci = np.sort(rs.choice(m, size=k, replace=False)) if k else np.zeros(0, dtype=np.int64)
    rs.shuffle(loci)
    beta_e = rs.normal(size=m_e + 1) * 0.2
    beta_g = rs.normal(size=k) * 0.3
    tau_g = rs.uniform(1.0, 8.0, size=k)
    return loci.astype(np.int64), beta_e, beta_g, tau_g


def oracle_yhat(bed, n, loci, beta_e, beta_g, E, miss=None):
    Ef = np.concatenate([np.ones((n, 1)), E], axis=1)
    ye = Ef @ beta_e
    yg = np.zeros(n)
    for l, b in zip(loci, beta_g):
        if miss is None:
            x = cpu.decode_column(bed, n, int(l), 0)
            x[x < 0] = 0
        else:
            off, idx, val = miss
            x = cpu.decode_column_overlay(bed, n, int(l), 0, idx[off[l]:off[l + 1]], val[off[l]:off[l + 1]])
        yg += b * x
    return ye, yg


@pytest.mark.parametrize("variant", [2, 1, 0])
@pytest.mark.parametrize("n,m,k,miss,tau_mode", [
    (30, 7, 2, 0.0, 0), (203, 300, 3, 0.01, 1), (1000, 500, 0, 0.0, 0), (1000, 500, 5, 0.0, 1),
    (5000, 1000, 4, 0.0, 0), (5000, 333, 3, 0.003, 1), (10007, 120, 2, 0.0, 0), (50000, 64, 3, 0.0, 1),
    (17, 3, 1, 0.0, 0)])
def test_scan_matches_oracle(api, variant, n, m, k, miss, tau_mode):
    payload, y, E = make_data(n, m, seed=3 * n + m, miss_rate=miss)
    bed, _ = oracle_store(payload, n, m, True)
    st = api.GenotypeStore(payload, n, m, recode_to_minor=True)
    st.set_phenotype(y, E)
    ch = api.Chain(st)
    ch.set_scan_variant(variant)
    loci, beta_e, beta_g, tau_g = random_state(n, m, 2, min(k, m), seed=n)
    k = loci.size
    off, idx, _ = cpu.missing_index(bed, n, m)
    rs = np.random.default_rng(5)
    val = rs.integers(0, 3, size=idx.size).astype(np.int8)
    for j in range(m):
        if off[j + 1] > off[j]:
            ch.set_missing(j, val[off[j]:off[j + 1]])
    miss_t = (off, idx, val)
    ye, yg = oracle_yhat(bed, n, loci, beta_e, beta_g, E, miss_t)
    stats = ch.residual(loci, beta_e, beta_g)
    r = ch.get_residual()
    r_or = y - (yg + ye)
    assert np.allclose(r, r_or, rtol=0, atol=1e-13 * max(1.0, np.abs(r_or).max()))
    assert stats["sum_r"] == pytest.approx(r_or.sum(), abs=1e-9)
    assert stats["xbxb"] == pytest.approx(yg @ yg, rel=1e-12, abs=1e-12)
    assert stats["ebxb"] == pytest.approx((ye - y) @ yg, rel=1e-11, abs=1e-10)
    assert stats["sum_yhat2"] == pytest.approx(((ye + yg) ** 2).sum(), rel=1e-12)
    # scan
    lmp_add, lmp_rem = -3.1, -2.7   # the two tabulated model-prior changes (sampler.cpp:52-76); any values do
    tau = 3.7 if tau_mode == 0 else rs.uniform(0.5, 9.0, size=m)
    model_ind = -np.ones(m, dtype=np.int32)
    for i, l in enumerate(loci):
        model_ind[l] = i
    sigma2 = 0.8
    xx = cpu.moments(bed, n, m)
    p_or, dot_or, _ = cpu.scan_A(bed, n, m, xx, y, ye + yg, model_ind, beta_g, tau_g, tau, tau_mode, sigma2, lmp_add, lmp_rem,
                                 miss=miss_t, want_dot=True)
    p = ch.scan(loci, beta_g, tau_g, sigma2, lmp_add, lmp_rem, tau=tau)
    assert np.abs(p - p_or).max() < 1e-9, np.abs(p - p_or).max()
    # the raw reduction x_j'r (dense part; the oracle's dot for in-model SNPs includes beta x'x, so compare off-model)
    dots = ch.scan_dots()
    G = cpu.decode_matrix(bed, n, m, 0)
    G[G < 0] = 0
    ref_dense = G.T @ r_or
    scale = np.sqrt((G * G).sum(axis=0)) * np.linalg.norm(r_or) + 1e-300
    assert (np.abs(dots - ref_dense) / scale).max() < 1e-12
    ch.close()
    st.close()


def test_scan_reference_goldens(api, golden, golden_data_dir):
    """p_r of the unmodified reference for small_modelspace (tests/golden/ref_outputs.npz)."""
    raw = cpu.read_bed(os.path.join(golden_data_dir, "small_modelspace.bed"), 30, 7)
    _, y = load_small(golden_data_dir)
    st = api.GenotypeStore(raw, 30, 7, recode_to_minor=True)
    st.set_phenotype(y)
    ch = api.Chain(st)
    pp = golden["small_prior"]
    P = cpu.Prior.make(7, 2, 2)
    # empty model, residual = y  (src/tests/raoblackwellizer_tests.hpp:31-93): beta_e = 0 gives y_hat = 0
    ch.residual([], [0.0], [])
    p = ch.scan([], [], [], 0.7, P.log_add([0] * 5, 0, 0), 0.0, tau=pp[8])
    assert np.abs(p - golden["small0_scan_empty_resid_y"]).max() < 1e-12
    # two SNPs in the model
    ch.residual([2, 5], [0.05], [0.4, -0.3])
    assert np.allclose(y - ch.get_residual(), golden["small0_two_snps_yhat"], rtol=0, atol=1e-14)
    p = ch.scan([2, 5], [0.4, -0.3], [11.0, 7.5], 0.9, P.log_add([2, 0, 0, 0, 0], 2, 0), P.log_add([1, 0, 0, 0, 0], 1, 0),
                tau=pp[8])
    assert np.abs(p - golden["small0_scan_two_snps"]).max() < 1e-12
    ch.close()
    st.close()


def test_scan_device_tau_draws_have_the_prior_distribution(api):
    """tau_mode 2 (throughput mode): 1/(alpha2 * Inv-chi2(nu, s2)) per SNP, drawn on the device.
    With y = 1 and an empty model the centred statistic rx is exactly 0, so p_r can be inverted for tau in
    closed form:  logit p = -0.5 log((v + tau)/tau)  =>  tau = v / (exp(-2 logit p) - 1)."""
    n, m = 64, 200000
    payload, _, _ = make_data(n, m, seed=12)
    y = np.ones(n)
    st = api.GenotypeStore(payload, n, m, recode_to_minor=False)
    st.set_phenotype(y)
    ch = api.Chain(st)
    ch.residual([], [0.0], [])
    nu, s2, alpha2, sigma2 = 5.0, 0.05, 1.3, 1.0
    kw = dict(tau_mode=2, seed=99, nu_tau2=nu, s2_tau2=s2, alpha2=alpha2)
    p = ch.scan([], [], [], sigma2, 0.0, 0.0, counter=1, **kw)
    p2 = ch.scan([], [], [], sigma2, 0.0, 0.0, counter=1, **kw)
    assert np.array_equal(p, p2)            # counter-based: reproducible
    p3 = ch.scan([], [], [], sigma2, 0.0, 0.0, counter=2, **kw)
    assert not np.array_equal(p, p3)        # a new scan counter gives new draws
    v = st.moments()[:, 1]
    ok = v > 0.5
    logit = np.log(p[ok]) - np.log1p(-p[ok])
    tau = v[ok] / np.expm1(-2.0 * logit)
    g = nu * s2 * alpha2 * tau / 2.0        # the underlying Gamma(nu/2, 1) draw
    assert g.mean() == pytest.approx(nu / 2, rel=0.02)
    assert g.var() == pytest.approx(nu / 2, rel=0.05)
    assert np.mean(g < 1.0) == pytest.approx(0.150855, abs=0.01)   # P[Gamma(2.5) < 1]
    ch.close()
    st.close()


# --------------------------------------------------------------------------- epilogue + weights
def test_adapt_and_partial_cdf_and_device_sampler(api):
    n, m = 100, 1000
    payload, y, E = make_data(n, m, seed=42)
    st = api.GenotypeStore(payload, n, m, recode_to_minor=True)
    st.set_phenotype(y, E)
    ch = api.Chain(st)
    q_add_min, q_rem_min = 1.0 / (m - 20), 1.0 / 20
    ch.init_proposal_flat(20.0 / m, q_add_min, q_rem_min)
    assert np.array_equal(ch.get_array("p_proposal"), np.full(m, 20.0 / m))
    ch.residual([], np.zeros(3), [])
    p = ch.scan([], [], [], 1.0, -3.0, 0.0, tau=4.0)
    prop0 = ch.get_array("p_proposal")
    ch.adapt(True, 0, True, 1, q_add_min, q_rem_min)
    assert np.array_equal(ch.get_array("p_rao"), cpu.running_mean(np.zeros(m), p, 0))
    prop = cpu.running_mean(prop0, p, 1)
    assert np.array_equal(ch.get_array("p_proposal"), prop)
    qa, qr = cpu.proposal_weights(prop, q_add_min, q_rem_min)
    assert np.array_equal(ch.get_array("q_add"), qa) and np.array_equal(ch.get_array("q_rem"), qr)
    bs, a_sums, r_sums = ch.partial_cdf()
    order = cpu.inorder_permutation(m)
    want = np.add.reduceat(qa[order], np.arange(0, m, bs))
    assert np.allclose(a_sums, want, rtol=1e-14)
    assert a_sums.sum() == pytest.approx(qa.sum(), rel=1e-13)
    # device sampler == in-order CDF search of the oracle, with zeroing
    zeroed = np.zeros(m, dtype=np.uint8)
    rs = np.random.default_rng(1)
    for step in range(60):
        if step % 3 == 0:
            j = int(rs.integers(0, m))
            zeroed[j] ^= 1
            ch.set_zeroed(0, j, bool(zeroed[j]))
        u = float(rs.uniform())
        snp, tot = ch.sample(0, u)
        assert snp == cpu.dd_sample(qa, zeroed, order, u)
        assert tot == pytest.approx(cpu.dd_total(qa, zeroed), rel=1e-13)
    # dd_rem: everything zeroed except the model's SNPs (sampler.cpp:601-605)
    ch.fill_zeroed(1, True)
    zr = np.ones(m, dtype=np.uint8)
    for j in (3, 500, 999):
        ch.set_zeroed(1, j, False)
        zr[j] = 0
    for u in (0.0, 0.3, 0.6, 0.999999):
        snp, tot = ch.sample(1, u)
        assert snp == cpu.dd_sample(qr, zr, order, u)
    ch.close()
    st.close()


# ------------------------------------------------------------------------------ missing overlay (a3, f2)
@pytest.mark.parametrize("n,m,miss", [(203, 300, 0.05), (4097, 64, 0.01), (50, 20, 0.4)])
def test_bulk_overlay_upload_and_cell_gather_match_oracle(api, n, m, miss):
    """bmg_chain_set_missing_all (DataModel::sample_missing, data_model.cpp:78-90: all SNPs re-imputed at once) and
    bmg_chain_get_cells (the cells the Gibbs step reads as current_model->x(i_miss, col), sampler.cpp:304-449): bit-exact
    against the oracle's overlay decode."""
    payload, y, E = make_data(n, m, seed=3 * n + 1, miss_rate=miss)
    bed, _ = oracle_store(payload, n, m, True)
    st = api.GenotypeStore(payload, n, m, recode_to_minor=True)
    st.set_phenotype(y, E)
    ch = api.Chain(st)
    off, idx, _ = cpu.missing_index(bed, n, m)
    rs = np.random.default_rng(n)
    val = rs.integers(0, 3, size=idx.size).astype(np.int8)
    ch.set_missing_all(val)

    def col(j):
        return cpu.decode_column_overlay(bed, n, int(j), 0, idx[off[j]:off[j + 1]], val[off[j]:off[j + 1]])

    for j in rs.choice(m, size=5, replace=False):
        assert np.array_equal(ch.get_column(int(j)), col(j))
    loci = rs.choice(m, size=min(9, m), replace=False).astype(np.int64)
    rows = np.unique(np.concatenate([idx[off[j]:off[j + 1]] for j in loci] + [rs.integers(0, n, size=7)])).astype(np.int32)
    cells = ch.get_cells(loci, rows)
    assert cells.shape == (loci.size, rows.size)
    for li, j in enumerate(loci):
        assert np.array_equal(cells[li] & 3, col(j)[rows].astype(np.int8))
        assert np.array_equal((cells[li] & 4) != 0, np.isin(rows, idx[off[j]:off[j + 1]]))   # bit 2: missing call
    # a second upload replaces the first; a per-SNP upload then overrides one SNP only
    val2 = rs.integers(0, 3, size=idx.size).astype(np.int8)
    ch.set_missing_all(val2)
    j0 = int(loci[0])
    ch.set_missing(j0, val[off[j0]:off[j0 + 1]])
    val2[off[j0]:off[j0 + 1]] = val[off[j0]:off[j0 + 1]]
    val = val2
    cells = ch.get_cells(loci, rows)
    for li, j in enumerate(loci):
        assert np.array_equal(cells[li] & 3, col(j)[rows].astype(np.int8))
    with pytest.raises(RuntimeError):
        ch.set_missing_all(val[:-1])
    with pytest.raises(RuntimeError):
        ch.get_cells(loci, np.array([n], dtype=np.int32))
    ch.close()
    st.close()


def test_device_imputation_draws_follow_the_prior_and_skip_the_model(api):
    """bmg_chain_impute_from_prior (throughput mode of DataModel::sample_missing): cells of SNPs in the model keep their
    values; every other missing cell is drawn from its SNP's observed genotype frequencies."""
    n, m = 4000, 40
    payload, y, E = make_data(n, m, seed=77, miss_rate=0.25)
    bed, _ = oracle_store(payload, n, m, True)
    st = api.GenotypeStore(payload, n, m, recode_to_minor=True)
    st.set_phenotype(y, E)
    ch = api.Chain(st)
    off, idx, prior3 = cpu.missing_index(bed, n, m)
    prior3 = np.asarray(prior3).reshape(m, 3)
    keep = np.array([3, 17, 30], dtype=np.int64)
    marker = np.full(idx.size, 2, dtype=np.int8)
    ch.set_missing_all(marker)
    ch.impute_from_prior(keep[::-1].copy(), 1234, 1)
    rows_all = np.arange(n, dtype=np.int32)
    cells = ch.get_cells(np.arange(m, dtype=np.int64), rows_all)
    z_max = 0.0
    for j in range(m):
        v = cells[j, idx[off[j]:off[j + 1]]]
        assert (v & 4).all()
        v = v & 3
        if j in keep:
            assert (v == 2).all()
            continue
        p = np.diff(np.concatenate([[0.0], prior3[j]])) / prior3[j, 2]
        cnt = np.bincount(v, minlength=3)[:3]
        assert cnt.sum() == v.size and v.min() >= 0 and v.max() <= 2
        z = np.abs(cnt - v.size * p) / np.sqrt(np.maximum(v.size * p * (1 - p), 1.0))
        z_max = max(z_max, z.max())
    assert z_max < 5.0, z_max
    # another counter gives other draws; the same (seed, counter) gives the same draws
    ch.impute_from_prior(keep, 1234, 2)
    cells2 = ch.get_cells(np.arange(m, dtype=np.int64), rows_all)
    assert (cells2 != cells).any()
    ch.impute_from_prior(keep, 1234, 1)
    assert np.array_equal(ch.get_cells(np.arange(m, dtype=np.int64), rows_all), cells)
    ch.close()
    st.close()


# ------------------------------------------------------------------------------ scan with effect types (f1)
def test_typed_scan_matches_reference_goldens_and_oracle(api, tmp_path):
    """the scan for every configuration of effect types (A, H, D, R, AH), models holding SNPs of those
    types, shared and per-SNP prior precisions, with and without missing calls: p_r and the per-type distribution against
    the unmodified reference's outputs (tests/golden/ref_typed_scan.npz) and against the numpy restatement, to 1e-9."""
    from tests import typed_scan_cases as tc
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "ref_typed_scan.npz"))
    for tag, types, miss, indiv in tc.CASES:
        c = tc.load_case(g, tag, str(tmp_path / tag))
        st = api.GenotypeStore(c["payload"], tc.N, tc.M_G, recode_to_minor=True)
        st.set_phenotype(c["y"], c["E"][:, 1:])
        ch = api.Chain(st)
        if c["miss_vals"].size:
            ch.set_missing_all(c["miss_vals"])
        loci, tt, bg = tc.model_terms(c)
        ch.residual_types(loci, tt, c["beta_e"], bg)
        loci_type = np.array([c["codes"][int(ti)] for ti in c["model_ti"]], dtype=np.int32)
        p_r, prt = ch.scan_types(c["codes"], c["model_snp"], loci_type, c["model_beta"], c["model_tau"], c["sigma2"], c["lmp_add"],
                                 c["lmp_rem"], tau_shared=c["tau_shared"], tau_snp=c["tau_snp"] if c["indiv"] else None)
        assert np.allclose(p_r, c["p_r"], rtol=1e-9, atol=1e-12), tag
        o_pr, o_prt = tc.oracle_scan(c)
        assert np.allclose(p_r, o_pr, rtol=1e-9, atol=1e-12), tag
        if len(c["codes"]) > 1:
            assert np.allclose(prt, c["prt"], rtol=1e-9, atol=1e-12), tag
            assert np.allclose(prt, o_prt, rtol=1e-9, atol=1e-12), tag
        if tag == "A_H_D_R":   # the intended moment layout is available too, and differs from the reference's
            p2, _ = ch.scan_types(c["codes"], c["model_snp"], loci_type, c["model_beta"], c["model_tau"], c["sigma2"], c["lmp_add"],
                                  c["lmp_rem"], tau_shared=c["tau_shared"], tau_snp=c["tau_snp"] if c["indiv"] else None,
                                  reference_offsets=False)
            o2, _ = tc.oracle_scan(c, reference_offsets=False)
            assert np.allclose(p2, o2, rtol=1e-9, atol=1e-12) and np.abs(p2 - p_r).max() > 1e-6
        ch.close()
        st.close()


def test_typed_columns_feed_the_column_statistics(api):
    """Typed columns (H, D, R) are materialised as 0/1 packed columns by the overlay cache, so the fitted values of a typed
    model equal the oracle's: checked through the residual of bmg_chain_residual_types, with missing calls."""
    n, m = 1203, 50
    payload, y, E = make_data(n, m, seed=91, miss_rate=0.04)
    bed, _ = oracle_store(payload, n, m, True)
    st = api.GenotypeStore(payload, n, m, recode_to_minor=True)
    st.set_phenotype(y, E)
    ch = api.Chain(st)
    off, idx, _ = cpu.missing_index(bed, n, m)
    rs = np.random.default_rng(2)
    val = rs.integers(0, 3, size=idx.size).astype(np.int8)
    ch.set_missing_all(val)
    loci = np.array([4, 4, 9, 17, 30, 41], dtype=np.int64)
    tt = np.array([0, 1, 2, 3, 1, 0], dtype=np.int32)
    bg = rs.normal(size=loci.size)
    be = rs.normal(size=3)

    def col(j):
        return cpu.decode_column_overlay(bed, n, int(j), 0, idx[off[j]:off[j + 1]], val[off[j]:off[j + 1]])

    X = np.column_stack([np.ones(n), E] + [cpu.typed(col(j), int(t)) for j, t in zip(loci, tt)])
    for _ in range(2):   # second call: every typed column comes from the cache
        ch.residual_types(loci, tt, be, bg)
        assert np.allclose(ch.get_residual(), y - X @ np.concatenate([be, bg]), rtol=1e-12, atol=1e-10)
    # new imputed values for SNP 4: its cached additive AND heterozygous columns must follow
    j = 4
    val[off[j]:off[j + 1]] = (val[off[j]:off[j + 1]] + 1) % 3
    ch.set_missing(j, val[off[j]:off[j + 1]])
    X = np.column_stack([np.ones(n), E] + [cpu.typed(col(jj), int(t)) for jj, t in zip(loci, tt)])
    ch.residual_types(loci, tt, be, bg)
    assert np.allclose(ch.get_residual(), y - X @ np.concatenate([be, bg]), rtol=1e-12, atol=1e-10)
    ch.close()
    st.close()


# ------------------------------------------------------------------------------ column stats (a7)
@pytest.mark.parametrize("n,m,miss", [(203, 300, 0.02), (5000, 200, 0.0), (40000, 50, 0.001)])
def test_column_stats_match_oracle(api, n, m, miss):
    payload, y, E = make_data(n, m, seed=n + 1, miss_rate=miss)
    bed, _ = oracle_store(payload, n, m, True)
    st = api.GenotypeStore(payload, n, m, recode_to_minor=True)
    st.set_phenotype(y, E)
    ch = api.Chain(st)
    off, idx, _ = cpu.missing_index(bed, n, m)
    rs = np.random.default_rng(8)
    val = rs.integers(0, 3, size=idx.size).astype(np.int8)
    for j in range(m):
        if off[j + 1] > off[j]:
            ch.set_missing(j, val[off[j]:off[j + 1]])
    loci = rs.choice(m, size=6, replace=False).astype(np.int64)
    cand = np.array([c for c in rs.choice(m, size=12, replace=False) if c not in loci][:5], dtype=np.int64)

    def col(j):
        return cpu.decode_column_overlay(bed, n, int(j), 0, idx[off[j]:off[j + 1]], val[off[j]:off[j + 1]])

    Ef = np.concatenate([np.ones((n, 1)), E], axis=1)
    X = np.concatenate([Ef, np.stack([col(l) for l in loci], axis=1)], axis=1)
    xy, xe, xm, xc = ch.column_stats(cand, loci)
    for ci, c in enumerate(cand):
        x = col(c)
        xy_or, xxcol = cpu.column_stats(x, y, X)
        assert xy[ci] == pytest.approx(xy_or, rel=1e-12, abs=1e-10)
        assert np.allclose(xe[ci], xxcol[:3], rtol=1e-12, atol=1e-10)
        assert np.array_equal(xm[ci], xxcol[3:-1])         # genotype x genotype: exact integers
        assert xc[ci, ci] == xxcol[-1]
        for di, d in enumerate(cand):
            assert xc[ci, di] == col(d) @ x
    ch.close()
    st.close()


# ------------------------------------------------------------------------------------ probit (a11)
def test_probit_latent_update_matches_oracle(api):
    n, m = 4000, 50
    payload, y, E = make_data(n, m, seed=77)
    bed, _ = oracle_store(payload, n, m, True)
    st = api.GenotypeStore(payload, n, m, recode_to_minor=True)
    is_case = (y > 0).astype(np.uint8)
    st.set_phenotype(is_case.astype(np.float64), E)
    ch = api.Chain(st)
    loci, beta_e, beta_g, _ = random_state(n, m, 2, 3, seed=4)
    ch.residual(loci, beta_e, beta_g)
    ye, yg = oracle_yhat(bed, n, loci, beta_e, beta_g, E)
    rs = np.random.default_rng(2)
    u = rs.uniform(size=n)
    stats = ch.probit_update(is_case, u)
    z = ch.get_phenotype()
    z_or = cpu.probit_latent(ye + yg, is_case, u)
    assert np.allclose(z, z_or, rtol=1e-9, atol=1e-9)   # tolerance of the north star for fp64 statistics
    assert (z[is_case == 1] > 0).all() and (z[is_case == 0] <= 0).all()
    assert stats[0] == pytest.approx(z_or.sum(), rel=1e-9) and stats[1] == pytest.approx((z_or ** 2).sum(), rel=1e-9)
    # device-drawn uniforms: right support, reproducible per (seed, counter)
    ch.residual(loci, beta_e, beta_g)
    ch.probit_update(None, None, seed=5, counter=9)
    z1 = ch.get_phenotype()
    assert (z1[is_case == 1] > 0).all() and (z1[is_case == 0] <= 0).all()
    # the next residual is built from z
    ch.residual(loci, beta_e, beta_g)
    assert np.allclose(ch.get_residual(), z1 - (ye + yg), rtol=0, atol=1e-12)
    ch.close()
    st.close()


@pytest.mark.parametrize("n,m", [(5120, 20000), (5000, 20000), (3000, 50000), (8192, 50000), (50000, 20000)])
def test_scan_imma_repeated_launches_are_bit_identical(api, n, m):
    """The tensor-core scan accumulates in integers, so every launch on the same residual must give the same
    bits whatever the timing (launch after an idle GPU, cache-resident store, ragged last chunk): a regression
    guard for the producer/consumer stage ring.  The fp64 kernel is the cross-check."""
    import torch
    B = (n + 3) // 4
    g = torch.Generator(device="cuda").manual_seed(n + m)
    raw = torch.randint(0, 256, (m * B,), dtype=torch.uint8, device="cuda", generator=g)
    raw &= 0b10111011
    y = np.random.default_rng(n).normal(size=n)
    st = api.GenotypeStore(None, n, m, recode_to_minor=True, payload_device_ptr=raw.data_ptr())
    del raw
    st.set_phenotype(y)
    ch = api.Chain(st)
    ch.residual([], [0.0], [])
    ch.set_scan_variant(0)
    d0 = ch.scan_dots()
    ch.set_scan_variant(2)
    first = ch.scan_dots()
    assert np.abs(first - d0).max() <= 1e-12 * np.abs(d0).max()
    for _ in range(6):
        torch.cuda.synchronize()
        assert np.array_equal(ch.scan_dots(), first)
    ch.close(); st.close()


@pytest.mark.parametrize("k,m_c,server", [(30, 1, "1"), (150, 3, "1"), (150, 3, "0"), (240, 10, "1"), (270, 2, "1")])
def test_column_stats_large_models_all_paths(api, k, m_c, server, monkeypatch):
    """Model sizes across the fast path's range (<= 250 columns per request, second mailbox trip beyond 63) and the
    general path beyond it, through the persistent server and through a launch per request: exact integers for
    genotype x genotype, 1e-12 for x'y."""
    monkeypatch.setenv("BMG_COLSTATS_SERVER", server)
    n, m = 3000, 600
    payload, y, E = make_data(n, m, seed=77, miss_rate=0.0)
    bed, _ = oracle_store(payload, n, m, True)
    st = api.GenotypeStore(payload, n, m, recode_to_minor=True)
    st.set_phenotype(y, E)
    ch = api.Chain(st)
    rs = np.random.default_rng(k)
    perm = rs.permutation(m)
    loci, cand = perm[:k].astype(np.int64), perm[k:k + m_c].astype(np.int64)
    G = np.stack([cpu.decode_column(bed, n, int(j), 0) for j in np.concatenate([cand, loci])], axis=1)
    for rep in range(3):   # repeated requests: the server stays up between them
        xy, xe, xm, xc = ch.column_stats(cand, loci)
        assert np.allclose(xy, G[:, :m_c].T @ y, rtol=1e-12, atol=1e-10)
        assert np.array_equal(xm, G[:, :m_c].T @ G[:, m_c:])
        assert np.array_equal(xc, G[:, :m_c].T @ G[:, :m_c])
    ch.close()
    st.close()


# ------------------------------------------------------- full-size property checks (BASELINE C2)
def test_scan_full_size_linearity_and_checksum(api):
    """n=5000 x m=100000 (BASELINE config 2): size-independent properties instead of an oracle pass:
    linearity in the residual and a checksum  sum_j x_j'r = (sum_j x_j)'r  from the integer counts."""
    import torch
    n, m = 5000, 100000
    B = (n + 3) // 4
    g = torch.Generator(device="cuda").manual_seed(1)
    raw = torch.randint(0, 256, (m * B,), dtype=torch.uint8, device="cuda", generator=g)
    raw &= 0b10111011   # clear bit 2 of each nibble-pair => fewer 01 codes; remaining 01 are missing cells
    rs = np.random.default_rng(3)
    y = rs.normal(size=n)
    st = api.GenotypeStore(None, n, m, recode_to_minor=True, payload_device_ptr=raw.data_ptr())
    st.set_phenotype(y)
    ch = api.Chain(st)
    ch.residual([], [0.0], [])
    d1 = ch.scan_dots()
    ch.set_scan_variant(0)
    d0 = ch.scan_dots()
    assert np.abs(d1 - d0).max() <= 1e-12 * np.abs(d1).max()
    ch.set_scan_variant(2)
    d2 = ch.scan_dots()
    assert np.abs(d2 - d0).max() <= 1e-12 * np.abs(d1).max()
    assert np.array_equal(d2, ch.scan_dots())   # integer accumulation: bit-reproducible
    ch.set_scan_variant(1)
    # linearity: residual scaled by -2.5 (beta_e = 3.5 on an all-ones covariate gives y - 3.5; use y2 = a*y instead)
    st2 = api.GenotypeStore(None, n, m, recode_to_minor=True, payload_device_ptr=raw.data_ptr())
    st2.set_phenotype(-2.5 * y)
    ch2 = api.Chain(st2)
    ch2.residual([], [0.0], [])
    assert np.allclose(ch2.scan_dots(), -2.5 * d1, rtol=1e-13, atol=1e-9)
    # checksum against the per-individual allele totals computed independently with torch
    codes = torch.stack([(raw.view(m, B) >> s) & 3 for s in (0, 2, 4, 6)], dim=2).reshape(m, 4 * B)[:, :n]
    lut = torch.tensor([0, 0, 1, 2], dtype=torch.float64, device="cuda")   # 01 (missing) -> 0
    _, _, _, swapped = st.counts()
    sw = torch.from_numpy(swapped.astype(np.bool_)).cuda()
    vals = lut[codes.long()]
    vals = torch.where(sw[:, None] & (codes != 1), 2.0 - vals, vals)
    tot = vals.sum(dim=0).cpu().numpy()
    assert d1.sum() == pytest.approx(tot @ y, rel=1e-11)
    ch.close(); ch2.close(); st.close(); st2.close()


# ------------------------------------------------------- BASELINE's own shapes against the oracle (not only properties)
def _oracle_scan_case(api, n, m, n_model, seed):
    """Seeded Binomial(2, f_j) genotypes of a BASELINE shape generated on the device (bench.device_payload), the same bytes
    handed to the C oracle on the host: counts, moment cache, scan dot products and p_r for a model of n_model SNPs with
    per-SNP tau, and the column statistics of a few candidate SNPs."""
    import torch
    import bench
    from oracle import cpu
    payload_dev = bench.device_payload(n, 0, m, seed, torch.device("cuda", 0))
    payload = payload_dev.cpu().numpy()
    st = api.GenotypeStore(None, n, m, recode_to_minor=True, payload_device_ptr=payload_dev.data_ptr())
    del payload_dev
    rs = np.random.default_rng(seed)
    y = rs.normal(size=n)
    E = rs.uniform(size=(n, 2))
    st.set_phenotype(y, E)
    bed = payload.copy()
    cpu.recode_minor(bed, n, m)
    xx = cpu.moments(bed, n, m)
    assert np.array_equal(st.moments(), xx)
    loci = np.sort(rs.choice(m, size=n_model, replace=False)).astype(np.int64)
    beta_e = rs.normal(size=3) * 0.1
    beta_g = rs.normal(size=n_model) * 0.2
    tau_g = rs.uniform(0.5, 5.0, size=n_model)
    tau = rs.uniform(0.5, 5.0, size=m)
    ch = api.Chain(st)
    ch.residual(loci, beta_e, beta_g)
    yhat = beta_e[0] + E @ beta_e[1:]
    for l, b in zip(loci, beta_g):
        yhat = yhat + b * cpu.decode_column(bed, n, int(l), 0)
    model_ind = -np.ones(m, dtype=np.int32)
    model_ind[loci] = np.arange(n_model)
    want_p, want_dot, _ = cpu.scan_A(bed, n, m, xx, y, yhat, model_ind, beta_g, tau_g, tau, 1, 0.8, -8.5, -8.0, want_dot=True)
    got_p = ch.scan(loci, beta_g, tau_g, 0.8, -8.5, -8.0, tau=tau)
    got_dot = ch.scan_dots()
    scale = np.sqrt(n) * 2.0 * np.linalg.norm(y - yhat)   # >= ||x_j|| ||r||
    out_of_model = model_ind < 0   # for a SNP of the model the oracle reports the product with the residual WITHOUT that SNP
    assert np.abs(got_dot - want_dot)[out_of_model].max() <= 1e-12 * scale
    assert np.abs(got_p - want_p).max() <= 1e-9
    cand = rs.choice(m, size=3, replace=False).astype(np.int64)
    xy, xe, xm, xc = ch.column_stats(cand, loci)
    for i, c in enumerate(cand):
        x = cpu.decode_column(bed, n, int(c), 0)
        assert abs(xy[i] - x @ y) <= 1e-12 * np.linalg.norm(x) * np.linalg.norm(y) + 1e-9
        assert np.allclose(xe[i], np.concatenate([[x.sum()], x @ E]), rtol=1e-12, atol=1e-9)
        for q, l in enumerate(loci[:5]):
            assert xm[i, q] == x @ cpu.decode_column(bed, n, int(l), 0)
    ch.close()
    st.close()


def test_scan_and_column_stats_at_the_c2_shape_against_the_oracle(api):
    """BASELINE configs[1]: n = 5,000 x p = 100,000 in full ("full C2 p_r vs oracle.c, not just checksums")."""
    _oracle_scan_case(api, 5000, 100000, 20, 11)


def test_scan_and_column_stats_at_the_c4_shape_against_the_oracle(api):
    """BASELINE configs[3] shape: n = 50,000 individuals (3,125 packed words per column, 17 chunks of the tensor-core scan),
    10,000 SNPs of it -- what the oracle finishes in seconds."""
    _oracle_scan_case(api, 50000, 10000, 24, 12)


def test_probit_inverse_cdf_against_scipy_truncnorm_in_the_tails(api):
    """k_probit draws z ~ N(mu, 1) truncated to (0, inf) for cases and (-inf, 0] for controls by inverting the CDF through the
    lower tail.  Checked against implementations the repo does not own, for fitted values out to |mu| = 8 (where the
    admissible half-line holds 6e-16 of the mass) and uniforms out to 1e-12 from both ends:
      * exact quantiles from 50-digit arithmetic (mpmath: root of Phi(t) = u Phi(+-mu)), 1e-9 relative -- the north star's bound;
      * scipy.stats.truncnorm.isf / ppf, 1e-4 relative: scipy's own tail accuracy is 2e-6 .. 6e-5 on this grid (measured against
        the exact quantiles), so it only guards against gross errors."""
    from scipy.stats import truncnorm
    mus = np.concatenate([np.linspace(-8.0, 8.0, 33), [-6.5, 6.5]])
    us = np.concatenate([[1e-12, 1e-9, 1e-6, 1e-3], np.linspace(0.05, 0.95, 7), [1 - 1e-3, 1 - 1e-6, 1 - 1e-9, 1 - 1e-12]])
    MU, U = np.meshgrid(mus, us, indexing="ij")
    mu, u = np.concatenate([MU.ravel(), MU.ravel()]), np.concatenate([U.ravel(), U.ravel()])
    is_case = np.concatenate([np.ones(MU.size, dtype=np.uint8), np.zeros(MU.size, dtype=np.uint8)])
    n, m = mu.size, 16
    rs = np.random.default_rng(5)
    payload = rs.integers(0, 256, size=m * ((n + 3) // 4), dtype=np.uint8) & 0b10111011
    st = api.GenotypeStore(payload, n, m, recode_to_minor=False)
    st.set_phenotype(is_case.astype(np.float64), mu.reshape(n, 1))   # one covariate carrying the fitted values
    ch = api.Chain(st)
    ch.residual([], [0.0, 1.0], [])                                   # y_hat = 0 * 1 + 1 * mu
    ch.probit_update(is_case, u)
    z = ch.get_phenotype()
    ch.close()
    st.close()
    assert np.isfinite(z).all() and (z[is_case == 1] > 0).all() and (z[is_case == 0] <= 0).all()
    # cases are drawn through the survival function (u -> 0 is the far tail), controls through the CDF
    loose = np.where(is_case == 1, truncnorm.isf(u, -mu, np.inf, loc=mu), truncnorm.ppf(u, -np.inf, -mu, loc=mu))
    assert (np.abs(z - loose) / (np.abs(loose) + 1e-3)).max() < 1e-4
    mp = pytest.importorskip("mpmath")
    mp.mp.dps = 50
    exact = np.empty(n)
    for i in range(n):
        m_i, u_i = mp.mpf(float(mu[i])), mp.mpf(float(u[i]))
        if is_case[i]:   # P(Z > z) = u  <=>  Phi(mu - z) = u Phi(mu)
            t = _mp_norm_ppf(mp, u_i * mp.ncdf(m_i))
            exact[i] = float(m_i - t)
        else:            # P(Z <= z) = u  <=>  Phi(z - mu) = u Phi(-mu)
            t = _mp_norm_ppf(mp, u_i * mp.ncdf(-m_i))
            exact[i] = float(m_i + t)
    err = np.abs(z - exact) / (np.abs(exact) + 1e-3)
    assert err.max() < 1e-9, (float(err.max()), float(mu[err.argmax()]), float(u[err.argmax()]), float(z[err.argmax()]), float(exact[err.argmax()]))


def _mp_norm_ppf(mp, p):
    """Standard normal quantile in mpmath arithmetic: Newton iterations on Phi(t) = p from an asymptotic start."""
    if p > mp.mpf("0.5"):
        return -_mp_norm_ppf(mp, 1 - p)
    t = -mp.sqrt(-2 * mp.log(p)) if p < mp.mpf("0.1") else mp.mpf(-1)
    for _ in range(60):
        step = (mp.ncdf(t) - p) / mp.npdf(t)
        t -= step
        if abs(step) < mp.mpf(10) ** -40:
            break
    return t


def test_a_non_finite_residual_is_not_hidden_by_the_integer_scan(api):
    """The tensor-core scan quantises the residual; a NaN or an infinity in it must surface as NaN dot products (what the fp64
    variants give), not as finite-looking numbers."""
    n, m = 3000, 400
    payload, y, E = make_data(n, m, seed=21)
    y = y.copy()
    y[17] = np.nan
    st = api.GenotypeStore(payload, n, m, recode_to_minor=True)
    st.set_phenotype(y, E)
    ch = api.Chain(st)
    ch.residual([], np.zeros(E.shape[1] + 1), [])
    for variant in (2, 0):
        ch.set_scan_variant(variant)
        d = ch.scan_dots()
        assert np.isnan(d).all() if variant == 2 else np.isnan(d).any(), variant
    ch.close()
    st.close()


def test_one_whole_move_through_the_public_abi_only(api):
    """What INTEGRATION.md sections 4-6 tell a maintainer to do, done here with nothing but the exported entry points: an
    all-SNP scan, the proposal weights, one draw from the add distribution without replacement, the proposed SNP's column
    statistics, and the move's log acceptance ratio from a Gram matrix held on the host (src/sampler.hpp:899-927 for an
    addition of one SNP) -- every number against the oracle on decoded dense columns."""
    n, m = 1500, 400
    payload, y, E = make_data(n, m, seed=44)
    bed, _ = oracle_store(payload, n, m, True)
    xx = cpu.moments(bed, n, m)
    st = api.GenotypeStore(payload, n, m, recode_to_minor=True)
    st.set_phenotype(y, E)
    ch = api.Chain(st)
    D = np.concatenate([np.ones((n, 1)), E], axis=1)                     # the covariate block incl. the constant
    k_e = D.shape[1]
    loci = np.array([17, 230], dtype=np.int64)
    X = np.stack([cpu.decode_column(bed, n, int(j), 0) for j in loci], axis=1)
    tau_e = np.concatenate([[0.0], np.full(k_e - 1, 1.0)])
    tau_g = np.array([2.0, 3.5])
    nu, s2 = 1.0, 0.8
    yy = float(y @ y)
    # current model: Gram matrix from the ABI's column statistics of the model's own SNPs
    xy_l, xe_l, _, xc_l = ch.column_stats(loci, np.zeros(0, dtype=np.int64))
    G = np.zeros((k_e + 2, k_e + 2))
    G[:k_e, :k_e] = D.T @ D
    G[:k_e, k_e:] = xe_l.T
    G[k_e:, k_e:] = xc_l
    G = np.triu(G) + np.triu(G, 1).T
    rhs = np.concatenate([D.T @ y, xy_l])
    assert np.allclose(G[k_e:, k_e:], X.T @ X) and np.allclose(rhs[k_e:], X.T @ y, rtol=1e-12)
    ll_cur, U, v, S = cpu.log_marginal(G, np.concatenate([tau_e, tau_g]), rhs, nu * s2 + yy, n + nu)
    beta = np.linalg.solve(U, v)
    # scan and proposal weights on the device, one draw from the add distribution (model SNPs zeroed)
    ch.residual(loci, beta[:k_e], beta[k_e:])
    p_r = ch.scan(loci, beta[k_e:], tau_g, 0.9, -3.0, -2.5, tau=4.0)
    model_ind = -np.ones(m, dtype=np.int32)
    model_ind[loci] = [0, 1]
    yhat = D @ beta[:k_e] + X @ beta[k_e:]
    want = cpu.scan_A(bed, n, m, xx, y, yhat, model_ind, beta[k_e:], tau_g, 4.0, 0, 0.9, -3.0, -2.5)
    assert np.abs(p_r - want).max() < 1e-9
    q_add_min, q_rem_min = 1.0 / (m - 5.0), 1.0 / 5.0
    ch.init_proposal_flat(5.0 / m, q_add_min, q_rem_min)
    ch.adapt(0, 0, 1, 1, q_add_min, q_rem_min)                           # p_proposal = mean(flat, p_r); q_add, q_rem; partial CDFs
    for j in loci:
        ch.set_zeroed(0, int(j), True)
    u = 0.6180339887
    snp, total = ch.sample(0, u)
    p_prop = 0.5 * (5.0 / m) + 0.5 * want
    w = np.maximum(p_prop, q_add_min)
    zero = np.zeros(m, dtype=np.uint8)
    zero[loci] = 1
    order = cpu.inorder_permutation(m)
    assert total == pytest.approx(cpu.dd_total(w, zero), rel=1e-12)
    assert snp == cpu.dd_sample(w, zero, order, u)
    # the proposed SNP against y, E and the model: one call; the new factor column on the host
    xy_c, xe_c, xm_c, xc_c = ch.column_stats(np.array([snp], dtype=np.int64), loci)
    x_new = cpu.decode_column(bed, n, int(snp), 0)
    o_xy, o_col = cpu.column_stats(x_new, y, np.concatenate([D, X], axis=1))
    got_col = np.concatenate([xe_c[0], xm_c[0], [xc_c[0, 0]]])
    assert xy_c[0] == pytest.approx(o_xy, rel=1e-12) and np.allclose(got_col, o_col, rtol=1e-12, atol=1e-9)
    assert np.array_equal(xm_c[0], X.T @ x_new)                          # genotype x genotype products are exact integers
    tau_new = 1.7
    G2 = np.zeros((k_e + 3, k_e + 3))
    G2[:-1, :-1] = G
    G2[:-1, -1] = got_col[:-1]
    G2[-1, :-1] = got_col[:-1]
    G2[-1, -1] = got_col[-1]
    ll_new, _, _, _ = cpu.log_marginal(G2, np.concatenate([tau_e, tau_g, [tau_new]]), np.concatenate([rhs, xy_c]), nu * s2 + yy, n + nu)
    # ... which must be the likelihood of the model with the decoded column appended
    Xn = np.concatenate([D, X, x_new[:, None]], axis=1)
    ll_ref, _, _, _ = cpu.log_marginal(Xn.T @ Xn, np.concatenate([tau_e, tau_g, [tau_new]]), Xn.T @ y, nu * s2 + yy, n + nu)
    assert ll_new == pytest.approx(ll_ref, rel=1e-10)
    log_q_forward = np.log(w[snp]) - np.log(total)
    assert np.isfinite(ll_new - ll_cur - log_q_forward)
    ch.close()
    st.close()
