"""Python mirror of the reference interface for the hot path, over the C ABI.

Names follow the reference classes they stand in for:
  GenotypeStore  ~ bmagwa::Data + PrecomputedSNPCovariances (src/data.hpp:45-88,
                   src/precomputed_snp_covariances.hpp:46-56)
  Chain          ~ bmagwa::DataModel + RaoBlackwellizer + the per-chain arrays of Sampler
                   (src/data_model.hpp:101-140, src/sampler.hpp:73-74,201-258)
  Sampler        ~ bmagwa::Sampler driven by an INI file (src/sampler.cpp:551-880)
All arguments are HOST numpy arrays; every call goes through libbmagwa_b200.so.
"""
import ctypes as C
import os

import numpy as np

from . import _lib
from ._lib import ScanParams, check, f64p, i32p, i64p, i8p, u8p, vp

A, H, D, R = 0, 1, 2, 3


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _pf(a):
    return a.ctypes.data_as(f64p) if a is not None else None


def _pi(a):
    return a.ctypes.data_as(i64p) if a is not None else None


def read_bed_payload(path, n, m_g):
    """Data::read_g (src/data.cpp:245-273): header check, then exactly ceil(n/4)*m_g bytes."""
    B = (n + 3) // 4
    with open(path, "rb") as fh:
        head = fh.read(3)
        if len(head) < 3:
            raise RuntimeError("BED file could not be opened")
        if head[0] != 0x6C or head[1] != 0x1B:
            raise RuntimeError("BED file not recognised (magic number does not match)")
        if head[2] != 0x01:
            raise RuntimeError("BED file not in snp-major format")
        payload = np.fromfile(fh, dtype=np.uint8, count=B * m_g)
    if payload.size != B * m_g:
        raise RuntimeError("Reading the BED file failed")
    return payload


class GenotypeStore:
    def __init__(self, bed_payload, n, m_g, recode_to_minor=False, device=0, snp_lo=0, snp_hi=None,
                 payload_device_ptr=None, bed_path=None):
        self.L = _lib.lib()
        snp_hi = m_g if snp_hi is None else snp_hi
        self.n, self.m_g, self.lo, self.hi, self.m = n, m_g, snp_lo, snp_hi, snp_hi - snp_lo
        h = vp()
        if bed_path is not None:   # streamed straight from the PLINK file
            check(self.L.bmg_store_create_from_bed(os.fsencode(bed_path), n, m_g, snp_lo, snp_hi, int(recode_to_minor), device,
                                                   C.byref(h)))
        elif payload_device_ptr is not None:
            check(self.L.bmg_store_create(vp(payload_device_ptr), 1, n, m_g, snp_lo, snp_hi, int(recode_to_minor), device,
                                          C.byref(h)))
        else:
            bed = np.ascontiguousarray(bed_payload, dtype=np.uint8)
            B = (n + 3) // 4
            if bed.size != B * self.m:
                raise ValueError("payload must hold (snp_hi - snp_lo) * ceil(n/4) bytes")
            check(self.L.bmg_store_create(bed.ctypes.data_as(vp), 0, n, m_g, snp_lo, snp_hi, int(recode_to_minor), device,
                                          C.byref(h)))
        self.h = h
        self.m_e = 0

    @classmethod
    def from_ini(cls, ini_path, snp_lo=0, snp_hi=None, device=-1):
        """SNPs [snp_lo, snp_hi) of the data set an INI file names, phenotype and covariates set (bmg_store_create_from_ini)."""
        self = cls.__new__(cls)
        self.L = _lib.lib()
        h = vp()
        check(self.L.bmg_store_create_from_ini(os.fsencode(ini_path), snp_lo, -1 if snp_hi is None else snp_hi, device, C.byref(h)))
        self.h = h
        n, m_g, lo, hi, nmiss = (C.c_int64() for _ in range(5))
        m_e = C.c_int()
        check(self.L.bmg_store_dims(h, C.byref(n), C.byref(m_g), C.byref(lo), C.byref(hi), C.byref(m_e), C.byref(nmiss)))
        self.n, self.m_g, self.lo, self.hi, self.m, self.m_e = n.value, m_g.value, lo.value, hi.value, hi.value - lo.value, m_e.value
        return self

    def close(self):
        if getattr(self, "h", None):
            self.L.bmg_store_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_phenotype(self, y, covariates=None):
        """covariates: (n, m_e-1) without the constant; the column of ones is prepended (data.hpp:50,56)."""
        y = _f64(y)
        ones = np.ones((self.n, 1))
        E = ones if covariates is None or np.size(covariates) == 0 else np.concatenate([ones, _f64(covariates).reshape(self.n, -1)], axis=1)
        Ef = np.asfortranarray(E)
        self.m_e = Ef.shape[1]
        check(self.L.bmg_store_set_phenotype(self.h, _pf(y), _pf(Ef), self.m_e))

    def counts(self):
        n1 = np.zeros(self.m, dtype=np.int32)
        n2 = np.zeros(self.m, dtype=np.int32)
        nm = np.zeros(self.m, dtype=np.int32)
        sw = np.zeros(self.m, dtype=np.uint8)
        check(self.L.bmg_store_counts(self.h, n1.ctypes.data_as(i32p), n2.ctypes.data_as(i32p), nm.ctypes.data_as(i32p),
                                      sw.ctypes.data_as(u8p)))
        return n1, n2, nm, sw

    def summaries(self):
        out = np.zeros(6)
        check(self.L.bmg_store_summaries(self.h, _pf(out)))
        return dict(mean_x=out[0] / out[1] if out[1] else float("nan"), var_x=out[2] / out[3] if out[3] else float("nan"),
                    var_y=out[4], yy=out[5], raw=out)

    def moments(self):
        xx = np.zeros((self.m, 2))
        check(self.L.bmg_store_moments(self.h, _pf(xx)))
        return xx

    def missing(self):
        off = np.zeros(self.m + 1, dtype=np.int64)
        prior = np.zeros((self.m, 3))
        check(self.L.bmg_store_missing(self.h, _pi(off), None, _pf(prior)))
        idx = np.zeros(max(int(off[-1]), 1), dtype=np.int64)
        check(self.L.bmg_store_missing(self.h, _pi(off), _pi(idx), None))
        return off, idx[: int(off[-1])], prior

    def get_column(self, snp, type_=A):
        out = np.zeros(self.n)
        check(self.L.bmg_store_get_column(self.h, snp, type_, _pf(out)))
        return out

    def export(self):
        handle = np.zeros(64, dtype=np.uint8)
        wps = C.c_int64()
        check(self.L.bmg_store_export(self.h, handle.ctypes.data_as(vp), C.byref(wps)))
        return handle.tobytes(), wps.value

    def attach_peer(self, ipc_handle_bytes, snp_lo, snp_hi):
        buf = np.frombuffer(ipc_handle_bytes, dtype=np.uint8).copy()
        check(self.L.bmg_store_attach_peer(self.h, buf.ctypes.data_as(vp), 0, snp_lo, snp_hi))


class Chain:
    def __init__(self, store):
        self.L = _lib.lib()
        self.store = store
        h = vp()
        check(self.L.bmg_chain_create(store.h, C.byref(h)))
        self.h = h

    def close(self):
        if getattr(self, "h", None):
            self.L.bmg_chain_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def sync(self):
        check(self.L.bmg_chain_sync(self.h))

    def stream(self):
        return self.L.bmg_chain_stream(self.h)

    def set_missing(self, snp, vals):
        v = np.ascontiguousarray(vals, dtype=np.int8)
        check(self.L.bmg_chain_set_missing(self.h, snp, v.ctypes.data_as(i8p), v.size))

    def set_missing_all(self, vals):
        v = np.ascontiguousarray(vals, dtype=np.int8)
        check(self.L.bmg_chain_set_missing_all(self.h, v.ctypes.data_as(i8p), v.size))

    def get_column(self, snp, type_=A):
        out = np.zeros(self.store.n)
        check(self.L.bmg_chain_get_column(self.h, snp, type_, _pf(out)))
        return out

    def impute_from_prior(self, loci, seed, counter):
        loci = np.ascontiguousarray(loci, dtype=np.int64)
        check(self.L.bmg_chain_impute_from_prior(self.h, _pi(loci), loci.size, seed, counter))

    def get_cells(self, loci, rows):
        loci = np.ascontiguousarray(loci, dtype=np.int64)
        rows = np.ascontiguousarray(rows, dtype=np.int32)
        out = np.zeros((loci.size, rows.size), dtype=np.int8)
        check(self.L.bmg_chain_get_cells(self.h, _pi(loci), loci.size, rows.ctypes.data_as(i32p), rows.size,
                                         out.ctypes.data_as(i8p)))
        return out

    def residual(self, loci, beta_e, beta_g):
        loci = np.ascontiguousarray(loci, dtype=np.int64)
        be, bg = _f64(beta_e), _f64(beta_g)
        stats = np.zeros(9)
        check(self.L.bmg_chain_residual(self.h, _pi(loci), _pf(be), _pf(bg), loci.size, _pf(stats)))
        keys = ("sum_r", "sum_e", "sum_e2", "sum_g", "sum_g2", "sum_yhat", "sum_yhat2", "xbxb", "ebxb")
        return dict(zip(keys, stats))

    def residual_types(self, loci, term_types, beta_e, beta_g):
        """Fitted values / residual of a model of typed terms (an AH SNP = two terms, types 0 and 1)."""
        loci = np.ascontiguousarray(loci, dtype=np.int64)
        tt = np.ascontiguousarray(term_types, dtype=np.int32)
        be, bg = _f64(beta_e), _f64(beta_g)
        stats = np.zeros(9)
        check(self.L.bmg_chain_residual_types(self.h, _pi(loci), tt.ctypes.data_as(i32p), _pf(be), _pf(bg), loci.size, _pf(stats)))
        return stats

    def scan_types(self, types, loci, loci_type, beta2, tau2, sigma2, lmp_add, lmp_rem, tau_shared=None, tau_snp=None,
                   reference_offsets=True, fetch=True):
        """(p_r, p_r_types or None): bmg_chain_scan_types.  beta2 / tau2: (k, 2)."""
        from ._lib import ScanTypesParams
        prm = ScanTypesParams()
        prm.sigma2 = sigma2
        prm.n_types = len(types)
        for i, t in enumerate(types):
            prm.types[i] = int(t)
        la, lr = _f64(lmp_add), _f64(np.asarray(lmp_rem).reshape(-1))
        for i in range(5):
            prm.lmp_add[i] = la[i]
        for i in range(25):
            prm.lmp_rem[i] = lr[i]
        keep = None
        if tau_snp is not None:
            keep = _f64(np.asarray(tau_snp).reshape(-1))
            prm.tau_mode = 1
            prm.tau_host = _pf(keep)
        else:
            prm.tau_mode = 0
            ts = _f64(tau_shared)
            for i in range(4):
                prm.tau_shared[i] = ts[i]
        prm.reference_offsets = 1 if reference_offsets else 0
        loci = np.ascontiguousarray(loci, dtype=np.int64)
        lt = np.ascontiguousarray(loci_type, dtype=np.int32)
        b2 = _f64(np.asarray(beta2, dtype=np.float64).reshape(-1)) if loci.size else np.zeros(2)
        t2 = _f64(np.asarray(tau2, dtype=np.float64).reshape(-1)) if loci.size else np.zeros(2)
        m = self.store.m
        p_r = np.zeros(m) if fetch else None
        prt = np.zeros((m, len(types))) if len(types) > 1 and fetch else None
        check(self.L.bmg_chain_scan_types(self.h, _pi(loci), lt.ctypes.data_as(i32p), _pf(b2), _pf(t2), loci.size, C.byref(prm),
                                          _pf(p_r) if p_r is not None else None, _pf(prt) if prt is not None else None))
        return p_r, prt

    def get_residual(self):
        r = np.zeros(self.store.n)
        check(self.L.bmg_chain_get_residual(self.h, _pf(r)))
        return r

    def scan(self, loci, beta_g, tau_g, sigma2, lmp_add, lmp_rem, tau=None, tau_mode=None, seed=0, counter=0,
             nu_tau2=0.0, s2_tau2=0.0, alpha2=1.0, fetch=True):
        """p_raoblackwell over the local shard.  tau: scalar (shared) or array of per-SNP values."""
        loci = np.ascontiguousarray(loci, dtype=np.int64)
        bg, tg = _f64(beta_g), _f64(tau_g)
        prm = ScanParams()
        prm.sigma2, prm.lmp_add, prm.lmp_rem = sigma2, lmp_add, lmp_rem
        keep = None
        if tau_mode is None:
            tau_mode = 0 if np.ndim(tau) == 0 else 1
        prm.tau_mode = tau_mode
        if tau_mode == 0:
            prm.tau_shared = float(tau)
        elif tau_mode == 1:
            keep = _f64(tau)
            assert keep.size == self.store.m
            prm.tau_host = _pf(keep)
        prm.tau_seed, prm.tau_counter, prm.nu_tau2, prm.s2_tau2, prm.alpha2 = seed, counter, nu_tau2, s2_tau2, alpha2
        out = np.zeros(self.store.m) if fetch else None
        check(self.L.bmg_chain_scan(self.h, _pi(loci), _pf(bg), _pf(tg), loci.size, C.byref(prm), _pf(out)))
        return out

    def scan_dots(self, fetch=True):
        out = np.zeros(self.store.m) if fetch else None
        check(self.L.bmg_chain_scan_dots(self.h, _pf(out)))
        return out

    def set_scan_variant(self, v):
        check(self.L.bmg_chain_set_scan_variant(self.h, v))

    def adapt(self, update_rao, n_rao_mean, update_proposal, n_prop_mean, q_add_min, q_rem_min):
        check(self.L.bmg_chain_adapt(self.h, int(update_rao), n_rao_mean, int(update_proposal), n_prop_mean, q_add_min,
                                     q_rem_min))

    def init_proposal_flat(self, value, q_add_min, q_rem_min):
        check(self.L.bmg_chain_init_proposal_flat(self.h, value, q_add_min, q_rem_min))

    def get_array(self, name):
        which = dict(p_r=0, p_rao=1, p_proposal=2, q_add=3, q_rem=4)[name]
        out = np.zeros(self.store.m)
        check(self.L.bmg_chain_get_array(self.h, which, _pf(out)))
        return out

    def partial_cdf(self):
        nb, bs = C.c_int64(), C.c_int64()
        check(self.L.bmg_chain_partial_cdf(self.h, C.byref(nb), C.byref(bs), None, None))
        a, r = np.zeros(nb.value), np.zeros(nb.value)
        check(self.L.bmg_chain_partial_cdf(self.h, C.byref(nb), C.byref(bs), _pf(a), _pf(r)))
        return bs.value, a, r

    def sample(self, which, u01):
        snp, tot = C.c_int64(), C.c_double()
        check(self.L.bmg_chain_sample(self.h, which, u01, C.byref(snp), C.byref(tot)))
        return snp.value, tot.value

    def set_zeroed(self, which, snp, zeroed=True):
        check(self.L.bmg_chain_set_zeroed(self.h, which, snp, int(zeroed)))

    def fill_zeroed(self, which, zeroed=True):
        check(self.L.bmg_chain_fill_zeroed(self.h, which, int(zeroed)))

    def column_stats(self, cand, loci):
        cand = np.ascontiguousarray(cand, dtype=np.int64)
        loci = np.ascontiguousarray(loci, dtype=np.int64)
        m_c, k, m_e = cand.size, loci.size, self.store.m_e
        xy = np.zeros(m_c)
        xe = np.zeros((m_c, m_e))
        xm = np.zeros((m_c, max(k, 1)))
        xc = np.zeros((m_c, m_c))
        check(self.L.bmg_chain_column_stats(self.h, _pi(cand), m_c, _pi(loci), k, _pf(xy), _pf(xe), _pf(xm), _pf(xc)))
        return xy, xe, xm[:, :k], xc

    def probit_update(self, is_case=None, u01=None, seed=0, counter=0):
        ic = None if is_case is None else np.ascontiguousarray(is_case, dtype=np.uint8)
        u = None if u01 is None else _f64(u01)
        stats = np.zeros(2)
        check(self.L.bmg_chain_probit_update(self.h, None if ic is None else ic.ctypes.data_as(u8p), _pf(u), seed, counter,
                                             _pf(stats)))
        return stats

    def get_phenotype(self):
        y = np.zeros(self.store.n)
        check(self.L.bmg_chain_get_phenotype(self.h, _pf(y)))
        return y


class Sampler:
    """The MH driver (src/sampler.cpp:551-880) behind bmg_sampler_*; reads the reference's INI file."""

    def __init__(self, ini_path, chain_index=0, device=0, store=None, comm=None, group=None, **options):
        self.L = _lib.lib()
        h = vp()
        self._comm = comm   # keeps the callback object alive as long as the sampler
        self._group = group
        if group is not None and comm is None:   # one of several chains over the sharded store (bmg_group_create)
            check(self.L.bmg_sampler_create_grouped(str(ini_path).encode(), chain_index, store.h, group.h, C.byref(h)))
        elif comm is not None:
            check(self.L.bmg_sampler_create_sharded(str(ini_path).encode(), chain_index, store.h, C.byref(comm.struct), C.byref(h)))
        elif store is None:
            check(self.L.bmg_sampler_create(str(ini_path).encode(), chain_index, device, C.byref(h)))
        else:
            check(self.L.bmg_sampler_create_on_store(str(ini_path).encode(), chain_index, store.h, C.byref(h)))
        self.h = h
        for k, v in options.items():
            self.set_option(k, v)

    def set_option(self, key, value):
        check(self.L.bmg_sampler_set_option(self.h, str(key).encode(), str(value).encode()))

    def begin(self):
        check(self.L.bmg_sampler_begin(self.h))

    def run(self, n_iter):
        check(self.L.bmg_sampler_run(self.h, n_iter))

    def end(self):
        check(self.L.bmg_sampler_end(self.h))

    def stats(self):
        out = np.zeros(8)
        check(self.L.bmg_sampler_stats(self.h, _pf(out)))
        keys = ("iterations", "accepted", "model_size", "log_likelihood", "move_seconds", "scan_seconds", "scans", "column_stats_seconds")
        return dict(zip(keys, out))

    def counters(self):
        out = np.zeros(12)
        check(self.L.bmg_sampler_counters(self.h, _pf(out), 12))
        keys = ("moves_with_additions", "served_from_memo", "partly_from_memo", "device_requests", "memo_pairs", "server_fallbacks",
                "delayed_rejection_seconds", "delayed_rejection_events", "epilogue_seconds", "gibbs_seconds", "probit_sweeps")
        return dict(zip(keys, out))

    def inclusion_counts(self, m_g):
        """(counts[m_g] uint32, number of thinned samples counted): running MCMC inclusion counts of the chain."""
        counts = np.zeros(m_g, dtype=np.uint32)
        ns = C.c_int64(0)
        check(self.L.bmg_sampler_inclusion_counts(self.h, counts.ctypes.data_as(C.POINTER(C.c_uint32)), C.byref(ns)))
        return counts, int(ns.value)

    def chain_stream(self):
        return self.L.bmg_chain_stream(self.L.bmg_sampler_chain(self.h))

    def close(self):
        if getattr(self, "h", None):
            self.L.bmg_sampler_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def ini_lookup(ini_path, section, key, default=""):
    """A value of the INI file through the library's own parser (bmg_ini_lookup)."""
    buf = C.create_string_buffer(1024)
    check(_lib.lib().bmg_ini_lookup(os.fsencode(ini_path), section.encode(), key.encode(), default.encode(), buf, len(buf)))
    return buf.value.decode()
