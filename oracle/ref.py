"""ctypes bindings of oracle/_ref/libbmagwa_ref.so -- TEST INFRASTRUCTURE.

The library is the UNMODIFIED reference (/root/reference/src) compiled against
the shims in oracle/shim plus oracle/ref_driver.cpp.  `available()` is False
where it has not been built (it is built in the authoring container, travels to
the GPU box as a prebuilt file, and cannot be rebuilt there).
"""
import ctypes as C
import os

import numpy as np

from . import HERE

LIB_PATH = os.path.join(HERE, "_ref", "libbmagwa_ref.so")
BIN_PATH = os.path.join(HERE, "_ref", "bmagwa_ref")
_LIB = None


def available() -> bool:
    if not os.path.exists(LIB_PATH):
        return False
    try:
        lib()
        return True
    except OSError:
        return False


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(LIB_PATH)
        L.refd_last_error.restype = C.c_char_p
        L.refd_open.restype = C.c_void_p
        L.refd_open.argtypes = [C.c_char_p, C.c_int]
        for f in ("refd_get_genotype", "refd_prior_log_add", "refd_prior_log_rem", "refd_prior_log_model", "refd_prior_log_swi",
                  "refd_model_loglik", "refd_scan_time", "refd_run_chain", "refd_continue_chain", "refd_rng_u01", "refd_rng_normal",
                  "refd_rng_sinvchi2_1", "refd_rng_sinvchi2_2", "refd_dd_total", "refd_gammaln"):
            getattr(L, f).restype = C.c_double
        L.refd_missing.restype = C.c_long
        L.refd_dd_sample.restype = C.c_long
        L.refd_open_chain.restype = C.c_void_p
        L.refd_open_chain.argtypes = [C.c_void_p, C.c_int]
        L.refd_rng_new.restype = C.c_void_p
        L.refd_dd_new.restype = C.c_void_p
        L.refd_set_blas_threads(1)
        _LIB = L
    return _LIB


def _p(a, t=C.c_double):
    return a.ctypes.data_as(C.POINTER(t)) if a is not None else None


class Ref:
    """One reference chain context: Options + Data + moment cache + Sampler (main.cpp:47-76)."""

    def __init__(self, ini_path, chain_index=0, parent=None):
        self.L = lib()
        if parent is not None:
            self.h = self.L.refd_open_chain(parent.h, chain_index)
        else:
            self.h = self.L.refd_open(str(ini_path).encode(), chain_index)
        if not self.h:
            raise RuntimeError(self.L.refd_last_error().decode())
        self.h = C.c_void_p(self.h)
        n, m_g, m_e, nt = C.c_long(), C.c_long(), C.c_long(), C.c_long()
        self.L.refd_sizes(self.h, C.byref(n), C.byref(m_g), C.byref(m_e), C.byref(nt))
        self.n, self.m_g, self.m_e, self.n_types = n.value, m_g.value, m_e.value, nt.value

    def close(self):
        if self.h:
            self.L.refd_close(self.h)
            self.h = None

    # ---- data
    def data_stats(self):
        out = np.zeros(4)
        self.L.refd_data_stats(self.h, _p(out))
        return dict(var_y=out[0], var_x=out[1], mean_x=out[2], yy=out[3])

    def y(self):
        out = np.zeros(self.n)
        self.L.refd_y(self.h, _p(out))
        return out

    def e(self):
        out = np.zeros((self.n, self.m_e), order="F")
        self.L.refd_e(self.h, _p(out))
        return out

    def get_genotype(self, ind, snp):
        return self.L.refd_get_genotype(self.h, C.c_long(ind), C.c_long(snp))

    def get_column(self, snp, type_=0, overlay=False):
        out = np.zeros(self.n)
        self.L.refd_get_column(self.h, C.c_long(snp), C.c_int(type_), C.c_int(int(overlay)), _p(out))
        return out

    def missing(self, snp):
        cnt = self.L.refd_missing(self.h, C.c_long(snp), None, None)
        if cnt == 0:
            return np.zeros(0, dtype=np.int64), None
        idx = np.zeros(cnt, dtype=np.int64)
        prior = np.zeros(3)
        self.L.refd_missing(self.h, C.c_long(snp), _p(idx, C.c_long), _p(prior))
        return idx, prior

    def set_miss_val(self, snp, k, val):
        self.L.refd_set_miss_val(self.h, C.c_long(snp), C.c_long(k), C.c_int(val))

    def get_miss_val(self, snp):
        cnt = self.L.refd_missing(self.h, C.c_long(snp), None, None)
        return np.array([self.L.refd_get_miss_val(self.h, C.c_long(snp), C.c_long(k)) for k in range(cnt)], dtype=np.int8)

    def moments(self):
        off = self.L.refd_moments_offset(self.h)
        out = np.zeros(self.m_g * off)
        self.L.refd_moments(self.h, _p(out))
        return out.reshape(self.m_g, off)

    def update_prexx_cov(self, snp, pre_xx):
        pre = np.array(pre_xx, dtype=np.float64)
        self.L.refd_update_prexx_cov(self.h, C.c_long(snp), _p(pre))
        return pre

    # ---- prior
    def prior_log_add(self, Ns, L, type_=0):
        Ns = np.asarray(Ns, dtype=np.int32)
        return self.L.refd_prior_log_add(self.h, _p(Ns, C.c_int), C.c_int(L), C.c_int(type_))

    def prior_log_rem(self, Ns, L, type_=0):
        Ns = np.asarray(Ns, dtype=np.int32)
        return self.L.refd_prior_log_rem(self.h, _p(Ns, C.c_int), C.c_int(L), C.c_int(type_))

    def prior_log_swi(self, Ns, type_add, type_rem):
        Ns = np.ascontiguousarray(Ns, dtype=np.int32)
        return self.L.refd_prior_log_swi(self.h, _p(Ns, C.c_int), C.c_int(type_add), C.c_int(type_rem))

    def prior_log_model(self, Ns):
        Ns = np.asarray(Ns, dtype=np.int32)
        return self.L.refd_prior_log_model(self.h, _p(Ns, C.c_int))

    def prior_params(self):
        out = np.zeros(10)
        self.L.refd_prior_params(self.h, _p(out))
        keys = ("g_a", "g_b", "n_plus_nu", "nus2_plus_yy", "alpha", "s2_sigma2", "nu_tau2_A", "s2_tau2_A",
                "inv_tau2_alpha2_A", "e_g")
        return dict(zip(keys, out))

    def prior_terms(self):
        """(4, 3) array: per term type A,H,D,R the shared inv_tau2_alpha2 value, nu_tau2, s2_tau2 (NaN if not allowed)."""
        out = np.zeros(12)
        self.L.refd_prior_terms(self.h, _p(out))
        return out.reshape(4, 3)

    def prior_set_alpha(self, a):
        self.L.refd_prior_set_alpha(self.h, C.c_double(a))

    # ---- model
    def model_add(self, snp, inv_tau2_alpha2, t_ind=0):
        v = np.atleast_1d(np.asarray(inv_tau2_alpha2, dtype=np.float64))
        v = np.concatenate([v, [0.0]])
        self.L.refd_model_add(self.h, C.c_long(snp), C.c_int(t_ind), _p(v))

    def model_remove(self, model_ind):
        self.L.refd_model_remove(self.h, C.c_int(model_ind))

    def model_compute_loglik(self):
        self.L.refd_model_compute_loglik(self.h)

    def model_loglik(self):
        return self.L.refd_model_loglik(self.h)

    def model_size(self):
        return self.L.refd_model_size(self.h)

    def model_cols(self):
        return self.L.refd_model_cols(self.h)

    def model_get(self, what):
        k = self.model_cols()
        names = dict(xx=0, l=1, xy=2, v=3, inv_tau2_alpha2=4, beta=5, scalars=6)
        w = names[what]
        if w in (0, 1):
            out = np.zeros((k, k), order="F")
        elif w == 6:
            out = np.zeros(5)
        else:
            out = np.zeros(k)
        self.L.refd_model_get(self.h, C.c_int(w), _p(out))
        if w == 6:
            return dict(zip(("sigma2", "syx_plus_vs2", "log_det_invQ", "log_det_invQ_plus_xx", "log_likelihood"), out))
        return out

    def model_loci(self):
        out = np.zeros(max(self.model_size(), 1), dtype=np.uint32)
        self.L.refd_model_loci(self.h, _p(out, C.c_uint))
        return out[: self.model_size()]

    def model_set_beta_sigma2(self, beta, sigma2):
        b = np.ascontiguousarray(beta, dtype=np.float64)
        assert b.size == self.model_cols()
        self.L.refd_model_set_beta_sigma2(self.h, _p(b), C.c_double(sigma2))

    def model_sample_beta_sigma2(self):
        self.L.refd_model_sample_beta_sigma2(self.h)

    def model_compute_pve(self):
        pves = np.zeros(3)
        y_hat = np.zeros(self.n)
        self.L.refd_model_compute_pve(self.h, _p(pves), _p(y_hat))
        return pves, y_hat

    def sample_alpha_and_tau2(self):
        self.L.refd_sample_alpha_and_tau2(self.h)

    def sample_missing(self):
        self.L.refd_sample_missing(self.h)

    def sample_missing_from_prior(self):
        self.L.refd_sample_missing_from_prior(self.h)

    # ---- scan
    def scan(self, y_hat=None):
        p_r = np.zeros(self.m_g)
        yh = None if y_hat is None else np.ascontiguousarray(y_hat, dtype=np.float64)
        prt = np.zeros((self.m_g, self.n_types)) if self.n_types > 1 else None
        self.L.refd_scan(self.h, _p(yh), _p(p_r), _p(prt))
        return p_r if prt is None else (p_r, prt)

    def scan_time(self, reps=1):
        return self.L.refd_scan_time(self.h, C.c_int(reps))

    def run_chain(self):
        t = self.L.refd_run_chain(self.h)
        if t < 0:
            raise RuntimeError(self.L.refd_last_error().decode())
        return t

    def set_do_n_iter(self, n):
        self.L.refd_set_do_n_iter(self.h, C.c_long(n))

    def continue_chain(self):
        # reference limitation: only safe while the model is empty (see ref_driver.cpp)
        t = self.L.refd_continue_chain(self.h)
        if t < 0:
            raise RuntimeError(self.L.refd_last_error().decode())
        return t

    def print_prior(self):
        self.L.refd_print_prior(self.h)


class RefRng:
    def __init__(self, seed, nu=1.0):
        self.L = lib()
        self.h = C.c_void_p(self.L.refd_rng_new(C.c_uint(seed), C.c_double(nu)))

    def u01(self):
        return self.L.refd_rng_u01(self.h)

    def normal(self):
        return self.L.refd_rng_normal(self.h)

    def sinvchi2_fixed(self, s2):
        return self.L.refd_rng_sinvchi2_1(self.h, C.c_double(s2))

    def sinvchi2(self, nu, s2):
        return self.L.refd_rng_sinvchi2_2(self.h, C.c_double(nu), C.c_double(s2))

    def __del__(self):
        try:
            self.L.refd_rng_free(self.h)
        except Exception:
            pass


class RefDD:
    """The reference's DiscreteDistribution with its own Rand(seed, 1.0)."""

    def __init__(self, w, seed):
        self.L = lib()
        w = np.ascontiguousarray(w, dtype=np.float64)
        self.m = w.size
        self.h = C.c_void_p(self.L.refd_dd_new(_p(w), C.c_long(w.size), C.c_uint(seed)))

    def sample(self):
        return self.L.refd_dd_sample(self.h)

    def zero(self, i):
        self.L.refd_dd_zero(self.h, C.c_long(i))

    def unzero(self, i):
        self.L.refd_dd_unzero(self.h, C.c_long(i))

    def total(self):
        return self.L.refd_dd_total(self.h)

    def update(self, w):
        w = np.ascontiguousarray(w, dtype=np.float64)
        self.L.refd_dd_update(self.h, _p(w))

    def __del__(self):
        try:
            self.L.refd_dd_free(self.h)
        except Exception:
            pass


def chol(a):
    k = a.shape[0]
    w = np.asfortranarray(np.triu(a), dtype=np.float64).copy(order="F")
    ok = lib().refd_chol(_p(w), C.c_int(k))
    return bool(ok), np.triu(w)


def chol_downdate(u, rem):
    k = u.shape[0]
    w = np.asfortranarray(u, dtype=np.float64).copy(order="F")
    lib().refd_chol_downdate(_p(w), C.c_int(k), C.c_int(rem))
    flat = w.ravel(order="F")[: (k - 1) * (k - 1)]
    return np.triu(flat.reshape((k - 1, k - 1), order="F"))


def chol_swapadj(u, col, v=None):
    k = u.shape[0]
    w = np.asfortranarray(u, dtype=np.float64).copy(order="F")
    vv = None if v is None else np.array(v, dtype=np.float64)
    lib().refd_chol_swapadj(_p(w), C.c_int(k), C.c_int(col), _p(vv))
    return np.triu(w), vv


def geometric_cdf(maxsize, p):
    out = np.zeros(maxsize)
    lib().refd_geometric_cdf(C.c_int(maxsize), C.c_double(p), _p(out))
    return out
