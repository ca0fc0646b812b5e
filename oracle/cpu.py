"""ctypes bindings of oracle/liboracle.so (oracle.c) -- TEST INFRASTRUCTURE."""
import ctypes as C
import os

import numpy as np

from . import HERE, build

_LIB = None


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(HERE, "liboracle.so")
        if not os.path.exists(path):
            build(ref=False)
        _LIB = C.CDLL(path)
        _LIB.orc_get_genotype.restype = C.c_double
        _LIB.orc_var.restype = C.c_double
        _LIB.orc_dot.restype = C.c_double
        _LIB.orc_prior_log_add.restype = C.c_double
        _LIB.orc_prior_log_rem.restype = C.c_double
        _LIB.orc_prior_log_model.restype = C.c_double
        _LIB.orc_log_marginal.restype = C.c_double
        _LIB.orc_dd_total.restype = C.c_double
        _LIB.orc_dd_sample.restype = C.c_long
        _LIB.orc_sample_discrete.restype = C.c_long
        _LIB.orc_rng_sizeof.restype = C.c_size_t
        for f in ("orc_rng_u01", "orc_rng_normal", "orc_rng_gamma", "orc_rng_sinvchi2_fixed", "orc_rng_sinvchi2"):
            getattr(_LIB, f).restype = C.c_double
        _LIB.orc_bytes_per_snp.restype = C.c_long
    return _LIB


def _p(a, t=C.c_double):
    return a.ctypes.data_as(C.POINTER(t)) if a is not None else None


def _u8(a):
    return a.ctypes.data_as(C.POINTER(C.c_uint8)) if a is not None else None


def read_bed(path, n, m_g):
    """data.cpp:245-273: 3-byte header check then ceil(n/4)*m_g payload bytes."""
    raw = np.fromfile(path, dtype=np.uint8)
    if raw[0] != 0x6C or raw[1] != 0x1B:
        raise RuntimeError("BED file not recognised (magic number does not match)")
    if raw[2] != 0x01:
        raise RuntimeError("BED file not in snp-major format")
    B = (n + 3) // 4
    if raw.size - 3 < B * m_g:
        raise RuntimeError("Reading the BED file failed")
    return np.ascontiguousarray(raw[3:3 + B * m_g]).copy()


def decode_column(bed, n, snp, type_=0):
    out = np.empty(n)
    lib().orc_decode_column(_u8(bed), C.c_long(n), C.c_long(snp), C.c_int(type_), _p(out))
    return out


def decode_matrix(bed, n, m_g, type_=0):
    return np.stack([decode_column(bed, n, j, type_) for j in range(m_g)], axis=1)


def recode_minor(bed, n, m_g):
    sw = np.zeros(m_g, dtype=np.uint8)
    lib().orc_recode_minor(_u8(bed), C.c_long(n), C.c_long(m_g), _u8(sw))
    return sw


def missing_index(bed, n, m_g):
    off = np.zeros(m_g + 1, dtype=np.int64)
    prior = np.zeros((m_g, 3))
    lib().orc_missing_index(_u8(bed), C.c_long(n), C.c_long(m_g), _p(off, C.c_long), None, _p(prior))
    idx = np.zeros(max(int(off[-1]), 1), dtype=np.int64)
    lib().orc_missing_index(_u8(bed), C.c_long(n), C.c_long(m_g), _p(off, C.c_long), _p(idx, C.c_long), _p(prior))
    return off, idx[: int(off[-1])], prior


def g_var_and_mean(bed, n, m_g):
    mean, var = C.c_double(), C.c_double()
    lib().orc_g_var_and_mean(_u8(bed), C.c_long(n), C.c_long(m_g), C.byref(mean), C.byref(var))
    return mean.value, var.value


def var(v):
    v = np.ascontiguousarray(v, dtype=np.float64)
    return lib().orc_var(_p(v), C.c_long(v.size))


def decode_column_overlay(bed, n, snp, type_, miss_idx, miss_val):
    out = np.empty(n)
    mi = np.ascontiguousarray(miss_idx, dtype=np.int64)
    mv = np.ascontiguousarray(miss_val, dtype=np.int8)
    lib().orc_decode_column_overlay(_u8(bed), C.c_long(n), C.c_long(snp), C.c_int(type_), _p(mi, C.c_long),
                                    mv.ctypes.data_as(C.POINTER(C.c_int8)), C.c_long(mi.size), _p(out))
    return out


def moment_layout(allow_types):
    at = np.asarray(allow_types, dtype=np.int32)
    terms = np.zeros(4, dtype=np.int32)
    offt = np.zeros(4, dtype=np.int32)
    off = lib().orc_moment_layout(_p(at, C.c_int), _p(terms, C.c_int), _p(offt, C.c_int))
    return off, terms, offt


def moments(bed, n, m_g, allow_types=(1, 0, 0, 0, 0)):
    at = np.asarray(allow_types, dtype=np.int32)
    off, _, _ = moment_layout(at)
    xx = np.zeros(m_g * off)
    lib().orc_moments(_u8(bed), C.c_long(n), C.c_long(m_g), _p(at, C.c_int), _p(xx))
    return xx.reshape(m_g, off)


def update_moments_for_missing(allow_types, n, miss_val, pre_xx):
    at = np.asarray(allow_types, dtype=np.int32)
    mv = np.ascontiguousarray(miss_val, dtype=np.int8)
    pre = np.array(pre_xx, dtype=np.float64)
    lib().orc_update_moments_for_missing(_p(at, C.c_int), C.c_long(n), mv.ctypes.data_as(C.POINTER(C.c_int8)),
                                         C.c_long(mv.size), _p(pre))
    return pre


class Prior(C.Structure):
    _fields_ = [("g_a", C.c_double), ("g_b", C.c_double), ("m_g", C.c_double),
                ("types_prior", C.c_double * 5), ("types_prior_sum", C.c_double)]

    @classmethod
    def make(cls, m_g, e_qg, var_qg, types_prior=(1, 1, 1, 1, 1), allow_types=(1, 0, 0, 0, 0)):
        p = cls()
        tp = np.asarray(types_prior, dtype=np.float64)
        at = np.asarray(allow_types, dtype=np.int32)
        ok = lib().orc_prior_init(C.byref(p), C.c_double(m_g), C.c_double(e_qg), C.c_double(var_qg), _p(tp), _p(at, C.c_int))
        if not ok:
            raise RuntimeError("Invalid a or b.")
        p._allow = at
        return p

    def log_add(self, Ns, L, type_=0):
        Ns = np.asarray(Ns, dtype=np.int32)
        return lib().orc_prior_log_add(C.byref(self), _p(Ns, C.c_int), C.c_int(L), C.c_int(type_))

    def log_rem(self, Ns, L, type_=0):
        Ns = np.asarray(Ns, dtype=np.int32)
        return lib().orc_prior_log_rem(C.byref(self), _p(Ns, C.c_int), C.c_int(L), C.c_int(type_))

    def log_model(self, Ns):
        Ns = np.asarray(Ns, dtype=np.int32)
        return lib().orc_prior_log_model(C.byref(self), _p(Ns, C.c_int), _p(self._allow, C.c_int))


def scan_A(bed, n, m_g, xx, y, y_hat, model_ind, beta_g, tau_g, tau, individual, sigma2, lmp_add, lmp_rem,
           miss=None, want_dot=False):
    """orc_scan_A; `tau` is a scalar (shared) or an array of m_g per-SNP values."""
    xx = np.ascontiguousarray(xx, dtype=np.float64)
    y = np.ascontiguousarray(y, dtype=np.float64)
    y_hat = np.ascontiguousarray(y_hat, dtype=np.float64)
    mi = np.ascontiguousarray(model_ind, dtype=np.int32)
    bg = np.ascontiguousarray(beta_g if len(beta_g) else [0.0], dtype=np.float64)
    tg = np.ascontiguousarray(tau_g if len(tau_g) else [0.0], dtype=np.float64)
    tau_arr = np.atleast_1d(np.asarray(tau, dtype=np.float64))
    stride = 1 if tau_arr.size == m_g and m_g > 1 else 0
    p_r = np.empty(m_g)
    dot = np.empty(m_g)
    rx = np.empty(m_g)
    if miss is not None:
        off, idx, val = miss
        off = np.ascontiguousarray(off, dtype=np.int64)
        idx = np.ascontiguousarray(idx, dtype=np.int64)
        val = np.ascontiguousarray(val, dtype=np.int8)
        a_off, a_idx, a_val = _p(off, C.c_long), _p(idx, C.c_long), val.ctypes.data_as(C.POINTER(C.c_int8))
    else:
        a_off = a_idx = a_val = None
    lib().orc_scan_A(_u8(bed), C.c_long(n), C.c_long(m_g), a_off, a_idx, a_val, _p(xx), _p(y), _p(y_hat),
                     _p(mi, C.c_int), _p(bg), _p(tg), _p(tau_arr), C.c_int(stride), C.c_int(int(individual)),
                     C.c_double(sigma2), C.c_double(lmp_add), C.c_double(lmp_rem), _p(p_r), _p(dot), _p(rx))
    return (p_r, dot, rx) if want_dot else p_r


def running_mean(mean, p_r, n_mean):
    mean = np.array(mean, dtype=np.float64)
    p_r = np.ascontiguousarray(p_r, dtype=np.float64)
    lib().orc_running_mean(_p(mean), _p(p_r), C.c_long(mean.size), C.c_long(n_mean))
    return mean


def proposal_weights(p_proposal, q_add_min, q_rem_min):
    p = np.ascontiguousarray(p_proposal, dtype=np.float64)
    qa, qr = np.empty_like(p), np.empty_like(p)
    lib().orc_proposal_weights(_p(p), C.c_long(p.size), C.c_double(q_add_min), C.c_double(q_rem_min), _p(qa), _p(qr))
    return qa, qr


def column_stats(x_new, y, X):
    x_new = np.ascontiguousarray(x_new, dtype=np.float64)
    y = np.ascontiguousarray(y, dtype=np.float64)
    Xf = np.asfortranarray(X, dtype=np.float64)
    n, k = Xf.shape
    xy = C.c_double()
    col = np.empty(k + 1)
    lib().orc_column_stats(_p(x_new), _p(y), _p(Xf), C.c_long(n), C.c_long(k), C.byref(xy), _p(col))
    return xy.value, col


def chol_upper(a):
    u = np.asfortranarray(np.triu(a), dtype=np.float64).copy(order="F")
    k = u.shape[0]
    ok = lib().orc_chol_upper(_p(u), C.c_int(k), C.c_int(k))
    return bool(ok), np.triu(u)


def chol_append(u, newcol, tau):
    k = u.shape[0]
    big = np.zeros((k + 1, k + 1), order="F")
    big[:k, :k] = u
    nc = np.ascontiguousarray(newcol, dtype=np.float64)
    ok = lib().orc_chol_append(_p(big), C.c_int(k), C.c_int(k + 1), _p(nc), C.c_double(tau))
    return bool(ok), big


def chol_delete(u, rem):
    k = u.shape[0]
    w = np.asfortranarray(u, dtype=np.float64).copy(order="F")
    lib().orc_chol_delete(_p(w), C.c_int(k), C.c_int(k), C.c_int(rem))
    return np.triu(w[: k - 1, : k - 1])


def chol_swapadj(u, col, v=None):
    k = u.shape[0]
    w = np.asfortranarray(u, dtype=np.float64).copy(order="F")
    vv = None if v is None else np.array(v, dtype=np.float64)
    lib().orc_chol_swapadj(_p(w), C.c_int(k), C.c_int(k), C.c_int(col), _p(vv))
    return np.triu(w), vv


def log_marginal(xx, tau, xy, nus2_plus_yy, n_plus_nu):
    k = len(xy)
    xxf = np.asfortranarray(np.triu(xx), dtype=np.float64).copy(order="F")
    tau = np.ascontiguousarray(tau, dtype=np.float64)
    xy = np.ascontiguousarray(xy, dtype=np.float64)
    u = np.zeros((k, k), order="F")
    v = np.zeros(k)
    S = C.c_double()
    ll = lib().orc_log_marginal(_p(xxf), _p(tau), _p(xy), C.c_int(k), C.c_int(k), C.c_double(nus2_plus_yy),
                                C.c_double(n_plus_nu), _p(u), _p(v), C.byref(S))
    return ll, np.triu(u), v, S.value


def pve(y_hat_e, y_hat_g, sigma2, have_e, have_g):
    e = np.ascontiguousarray(y_hat_e, dtype=np.float64)
    g = np.ascontiguousarray(y_hat_g, dtype=np.float64)
    out = np.zeros(3)
    lib().orc_pve(_p(e), _p(g), C.c_long(e.size), C.c_int(int(have_e)), C.c_int(int(have_g)), C.c_double(sigma2), _p(out))
    return out


def inorder_permutation(m):
    order = np.empty(m, dtype=np.int64)
    lib().orc_inorder_permutation(C.c_long(m), _p(order, C.c_long))
    return order


def dd_total(w, zeroed):
    w = np.ascontiguousarray(w, dtype=np.float64)
    z = np.ascontiguousarray(zeroed, dtype=np.uint8)
    return lib().orc_dd_total(_p(w), _u8(z), C.c_long(w.size))


def dd_sample(w, zeroed, order, u, total=None):
    w = np.ascontiguousarray(w, dtype=np.float64)
    z = np.ascontiguousarray(zeroed, dtype=np.uint8)
    order = np.ascontiguousarray(order, dtype=np.int64)
    if total is None:
        total = dd_total(w, z)
    return lib().orc_dd_sample(_p(w), _u8(z), _p(order, C.c_long), C.c_long(w.size), C.c_double(u), C.c_double(total))


def sample_discrete_naive(cumsum, u):
    c = np.ascontiguousarray(cumsum, dtype=np.float64)
    return lib().orc_sample_discrete_naive(_p(c), C.c_int(c.size), C.c_double(u))


def sample_discrete(cumsum, level, u):
    c = np.ascontiguousarray(cumsum, dtype=np.float64)
    return lib().orc_sample_discrete(_p(c), C.c_long(c.size), C.c_int(level), C.c_double(u))


def geometric_cdf(maxsize, p):
    out = np.empty(maxsize)
    lib().orc_geometric_cdf(C.c_int(maxsize), C.c_double(p), _p(out))
    return out


class Rng:
    """rand.hpp:36-191 restated (mt19937 + Boost.Random 1.47-1.55 variates)."""

    def __init__(self, seed, sinvchi2_nu=1.0):
        self._buf = C.create_string_buffer(lib().orc_rng_sizeof())
        lib().orc_rng_seed(self._buf, C.c_uint32(seed), C.c_double(sinvchi2_nu))

    def u01(self):
        return lib().orc_rng_u01(self._buf)

    def normal(self):
        return lib().orc_rng_normal(self._buf)

    def gamma(self, alpha):
        return lib().orc_rng_gamma(self._buf, C.c_double(alpha))

    def sinvchi2_fixed(self, s2):
        return lib().orc_rng_sinvchi2_fixed(self._buf, C.c_double(s2))

    def sinvchi2(self, nu, s2):
        return lib().orc_rng_sinvchi2(self._buf, C.c_double(nu), C.c_double(s2))


def probit_latent(mu, is_case, u):
    mu = np.ascontiguousarray(mu, dtype=np.float64)
    ic = np.ascontiguousarray(is_case, dtype=np.uint8)
    u = np.ascontiguousarray(u, dtype=np.float64)
    z = np.empty_like(mu)
    lib().orc_probit_latent(_p(mu), _u8(ic), _p(u), C.c_long(mu.size), _p(z))
    return z


# ---------------------------------------------------------------------------------------------------------------
# The all-SNP scan with several effect types (src/sampler.cpp:32-261 with n_types > 1, or one type other than A):
# numpy restatement, pinned against the unmodified reference in tests/test_oracle_goldens.py.
# ---------------------------------------------------------------------------------------------------------------
TYPE_A, TYPE_H, TYPE_D, TYPE_R, TYPE_AH = 0, 1, 2, 3, 4


def term_flags(types):
    """DataModel::allow_types_ / allow_terms_ (data_model.hpp:61-73)."""
    allow_types = [t in types for t in range(5)]
    allow_terms = [False] * 4
    for t in types:
        if t == TYPE_AH:
            allow_terms[TYPE_A] = allow_terms[TYPE_H] = True
        else:
            allow_terms[t] = True
    return allow_types, allow_terms


def typed(col, t):
    """DataModel::get_genotypes[t] on an additive column (data_model.cpp:41-72)."""
    if t == TYPE_A:
        return col.copy()
    if t == TYPE_H:
        return (col == 1).astype(np.float64)
    if t == TYPE_D:
        return (col > 0).astype(np.float64)
    return (col == 2).astype(np.float64)


def scan_types(columns0, columns, n, types, y, y_hat, model, sigma2, lmp_add, lmp_rem, tau_shared=None, tau_snp=None,
               reference_offsets=True):
    """p_r (m) and p_r_types (m x n_types; None when n_types == 1).
    columns0[j] / columns[j]: additive column of SNP j with missing cells = 0 / = the chain's imputed values.
    model: dict snp -> (type code, [beta1, beta2], [tau1, tau2]) for the SNPs in the model.
    tau_shared[4]: value per TERM type (shared-tau mode); tau_snp[m][n_terms]: per-SNP draws in increasing term code.
    reference_offsets: index the moment cache as the reference does -- offset_type[t] counts TERMS although a term
    occupies two slots (precomputed_snp_covariances.hpp:73-83 vs :113-118), so with several term types every non-first
    type reads (s, v) one slot early.  False: the intended layout."""
    m = len(columns)
    n_types = len(types)
    allow_types, allow_terms = term_flags(types)
    terms = [t for t in range(4) if allow_terms[t]]
    rank = {t: i for i, t in enumerate(terms)}
    offset = 2 * len(terms) + (1 if allow_types[TYPE_AH] else 0)
    r_cm = y - y_hat
    mr_cm = r_cm.sum() / n
    p_r = np.zeros(m)
    prt = np.zeros((m, n_types))
    for j in range(m):
        x0, x1 = columns0[j], columns[j]
        # PrecomputedSNPCovariances (missing = 0) then DataModel::update_prexx_cov (imputed values)
        pre = np.zeros(offset)
        for t in terms:
            xt = typed(x0, t)
            s0 = xt.sum()
            pre[2 * rank[t]] = s0
            pre[2 * rank[t] + 1] = xt @ xt - s0 * s0 / n
        if allow_types[TYPE_AH]:
            pre[offset - 1] = typed(x0, TYPE_A) @ typed(x0, TYPE_H) - typed(x0, TYPE_A).sum() * typed(x0, TYPE_H).sum() / n
        miss = np.flatnonzero(x0 != x1)
        # cells whose imputed value is 0 do not change anything either, so "x0 != x1" is enough
        if miss.size:
            sa_sh = pre[0] * pre[2] if allow_types[TYPE_AH] else 0.0
            for t in terms:
                vg = typed(x1, t)[miss]
                sv, sv2 = vg.sum(), (vg * vg).sum()
                old = pre[2 * rank[t]]
                pre[2 * rank[t]] += sv
                pre[2 * rank[t] + 1] += sv2 - sv * (old * 2.0 + sv) / n
            if allow_types[TYPE_AH]:
                pre[offset - 1] += (sa_sh - pre[0] * pre[2]) / n
                pre[offset - 1] += float((x1[miss] == 1).sum())
        # taus of this SNP
        tau = np.zeros(4)
        for t in terms:
            tau[t] = tau_shared[t] if tau_snp is None else tau_snp[j][rank[t]]
        if j in model:
            mt, betas, mtaus = model[j]
            if tau_snp is not None:
                if mt == TYPE_AH:
                    tau[TYPE_A], tau[TYPE_H] = mtaus[0], mtaus[1]
                else:
                    tau[mt] = mtaus[0]
            if mt == TYPE_AH:
                residual = r_cm + typed(x1, TYPE_A) * betas[0] + typed(x1, TYPE_H) * betas[1]
            else:
                residual = r_cm + typed(x1, mt) * betas[0]
            mr = residual.sum() / n
            lmp = lmp_rem[mt]
        else:
            residual, mr, lmp = r_cm, mr_cm, lmp_add
        p = np.zeros(n_types)
        for ti, t in enumerate(types):
            if t == TYPE_AH:
                sa, va = pre[0], pre[1] + tau[TYPE_A]
                sh, vh = pre[2], pre[3] + tau[TYPE_H]
                vah = pre[offset - 1]
                sum_log_q = -np.log(tau[TYPE_A]) - np.log(tau[TYPE_H])
                rxa = typed(x1, TYPE_A) @ residual - sa * mr
                rxh = typed(x1, TYPE_H) @ residual - sh * mr
                det = va * vh - vah * vah
                exp_term = (rxa * rxa * vh - 2.0 * rxa * rxh * vah + rxh * rxh * va) / det
            else:
                o = rank[t] if reference_offsets else 2 * rank[t]
                s = pre[o]
                det = pre[o + 1] + tau[t]
                sum_log_q = -np.log(tau[t])
                rx = typed(x1, t) @ residual - s * mr
                exp_term = rx * rx / det
            with np.errstate(invalid="ignore", divide="ignore"):
                p[ti] = exp_term / (2 * sigma2) - 0.5 * (np.log(det) + sum_log_q) + lmp[t]
        if n_types == 1:
            with np.errstate(over="ignore"):
                e = np.exp(p[0])
            p_r[j] = e / (1 + e) if np.isfinite(e) else 1.0
            continue
        mx = -np.inf
        for v in p:   # "if (p > max_types)": a NaN never becomes the maximum
            if v > mx:
                mx = v
        if not np.isfinite(mx):
            if mx > 0:
                w = np.array([1.0 if (not np.isfinite(v) and v > 0) else 0.0 for v in p])
                prt[j] = w / w.sum()
                p_r[j] = 1.0
            else:
                prt[j] = 1.0 / n_types
                p_r[j] = 0.0
            continue
        w = np.exp(p - mx)
        tot = 0.0
        for v in w:
            tot += v
        prt[j] = w / tot
        with np.errstate(over="ignore"):
            st = np.exp(np.log(tot) + mx)
        p_r[j] = st / (1 + st) if np.isfinite(st) else 1.0
    return p_r, (prt if n_types > 1 else None)
