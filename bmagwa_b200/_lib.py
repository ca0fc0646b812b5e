"""ctypes loader of libbmagwa_b200.so.  There is no Python/CPU fallback: if the library is
missing or cannot be loaded this raises, loudly."""
import ctypes as C
import os

PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(PKG, "libbmagwa_b200.so")

_LIB = None

i64, i32, f64, u64, u8p = C.c_int64, C.c_int32, C.c_double, C.c_uint64, C.POINTER(C.c_uint8)
f64p, i64p, i32p, i8p = C.POINTER(C.c_double), C.POINTER(C.c_int64), C.POINTER(C.c_int32), C.POINTER(C.c_int8)
vp = C.c_void_p


class ScanParams(C.Structure):
    """struct bmg_scan_params (include/bmagwa_b200.h)."""
    _fields_ = [("sigma2", f64), ("lmp_add", f64), ("lmp_rem", f64), ("tau_mode", i32), ("tau_shared", f64),
                ("tau_host", f64p), ("tau_seed", u64), ("tau_counter", u64), ("nu_tau2", f64), ("s2_tau2", f64),
                ("alpha2", f64)]


class ScanTypesParams(C.Structure):
    _fields_ = [("sigma2", f64), ("n_types", i32), ("types", i32 * 5), ("lmp_add", f64 * 5), ("lmp_rem", f64 * 25),
                ("tau_mode", i32), ("tau_shared", f64 * 4), ("tau_host", f64p), ("reference_offsets", i32)]


# name -> (restype, argtypes); every symbol declared in include/bmagwa_b200.h
SIGNATURES = {
    "bmg_abi_version": (C.c_int, []),
    "bmg_last_error": (C.c_char_p, []),
    "bmg_device_count": (C.c_int, []),
    "bmg_launch_count": (u64, []),
    "bmg_transfer_bytes": (None, [C.POINTER(u64), C.POINTER(u64)]),
    "bmg_store_create": (C.c_int, [vp, C.c_int, i64, i64, i64, i64, C.c_int, C.c_int, C.POINTER(vp)]),
    "bmg_store_create_from_bed": (C.c_int, [C.c_char_p, i64, i64, i64, i64, C.c_int, C.c_int, C.POINTER(vp)]),
    "bmg_store_destroy": (C.c_int, [vp]),
    "bmg_store_set_phenotype": (C.c_int, [vp, f64p, f64p, C.c_int]),
    "bmg_store_dims": (C.c_int, [vp, i64p, i64p, i64p, i64p, C.POINTER(C.c_int), i64p]),
    "bmg_store_counts": (C.c_int, [vp, i32p, i32p, i32p, u8p]),
    "bmg_store_summaries": (C.c_int, [vp, f64p]),
    "bmg_store_moments": (C.c_int, [vp, f64p]),
    "bmg_store_missing": (C.c_int, [vp, i64p, i64p, f64p]),
    "bmg_store_get_column": (C.c_int, [vp, i64, C.c_int, f64p]),
    "bmg_store_export": (C.c_int, [vp, vp, i64p]),
    "bmg_store_attach_peer": (C.c_int, [vp, vp, C.c_int, i64, i64]),
    "bmg_chain_create": (C.c_int, [vp, C.POINTER(vp)]),
    "bmg_chain_destroy": (C.c_int, [vp]),
    "bmg_chain_sync": (C.c_int, [vp]),
    "bmg_chain_stream": (vp, [vp]),
    "bmg_chain_set_missing": (C.c_int, [vp, i64, i8p, i64]),
    "bmg_chain_set_missing_all": (C.c_int, [vp, i8p, i64]),
    "bmg_chain_impute_from_prior": (C.c_int, [vp, i64p, C.c_int, u64, u64]),
    "bmg_chain_get_column": (C.c_int, [vp, i64, C.c_int, f64p]),
    "bmg_chain_get_cells": (C.c_int, [vp, i64p, C.c_int, i32p, i64, i8p]),
    "bmg_chain_residual": (C.c_int, [vp, i64p, f64p, f64p, C.c_int, f64p]),
    "bmg_chain_residual_types": (C.c_int, [vp, i64p, i32p, f64p, f64p, C.c_int, f64p]),
    "bmg_chain_scan_types": (C.c_int, [vp, i64p, i32p, f64p, f64p, C.c_int, C.POINTER(ScanTypesParams), f64p, f64p]),
    "bmg_chain_get_residual": (C.c_int, [vp, f64p]),
    "bmg_chain_scan": (C.c_int, [vp, i64p, f64p, f64p, C.c_int, C.POINTER(ScanParams), f64p]),
    "bmg_chain_scan_dots": (C.c_int, [vp, f64p]),
    "bmg_chain_set_scan_variant": (C.c_int, [vp, C.c_int]),
    "bmg_chain_scan_kernel_time": (C.c_int, [vp, C.c_int, f64p, i64p, C.c_int]),
    "bmg_chain_adapt": (C.c_int, [vp, C.c_int, i64, C.c_int, i64, f64, f64]),
    "bmg_chain_init_proposal_flat": (C.c_int, [vp, f64, f64, f64]),
    "bmg_chain_get_array": (C.c_int, [vp, C.c_int, f64p]),
    "bmg_chain_partial_cdf": (C.c_int, [vp, i64p, i64p, f64p, f64p]),
    "bmg_chain_sample": (C.c_int, [vp, C.c_int, f64, i64p, f64p]),
    "bmg_chain_set_zeroed": (C.c_int, [vp, C.c_int, i64, C.c_int]),
    "bmg_chain_fill_zeroed": (C.c_int, [vp, C.c_int, C.c_int]),
    "bmg_chain_column_stats": (C.c_int, [vp, i64p, C.c_int, i64p, C.c_int, f64p, f64p, f64p, f64p]),
    "bmg_chain_probit_update": (C.c_int, [vp, u8p, f64p, u64, u64, f64p]),
    "bmg_chain_get_phenotype": (C.c_int, [vp, f64p]),
    "bmg_sampler_create": (C.c_int, [C.c_char_p, C.c_int, C.c_int, C.POINTER(vp)]),
    "bmg_sampler_create_on_store": (C.c_int, [C.c_char_p, C.c_int, vp, C.POINTER(vp)]),
    "bmg_sampler_create_sharded": (C.c_int, [C.c_char_p, C.c_int, vp, vp, C.POINTER(vp)]),
    "bmg_sampler_set_option": (C.c_int, [vp, C.c_char_p, C.c_char_p]),
    "bmg_sampler_begin": (C.c_int, [vp]),
    "bmg_sampler_run": (C.c_int, [vp, i64]),
    "bmg_sampler_end": (C.c_int, [vp]),
    "bmg_sampler_stats": (C.c_int, [vp, f64p]),
    "bmg_sampler_counters": (C.c_int, [vp, f64p, C.c_int]),
    "bmg_ini_lookup": (C.c_int, [C.c_char_p, C.c_char_p, C.c_char_p, C.c_char_p, C.c_char_p, C.c_int]),
    "bmg_store_create_from_ini": (C.c_int, [C.c_char_p, C.c_int64, C.c_int64, C.c_int, C.POINTER(C.c_void_p)]),
    "bmg_group_create": (C.c_int, [vp, C.c_int, C.c_int, C.c_int, i64, C.c_char_p, C.POINTER(vp)]),
    "bmg_sampler_create_grouped": (C.c_int, [C.c_char_p, C.c_int, vp, vp, C.POINTER(vp)]),
    "bmg_group_serve": (C.c_int, [vp, i64]),
    "bmg_group_scan_chain": (vp, [vp]),
    "bmg_group_stats": (C.c_int, [vp, f64p]),
    "bmg_group_destroy": (C.c_int, [vp]),
    "bmg_sampler_inclusion_counts": (C.c_int, [vp, C.POINTER(C.c_uint32), i64p]),
    "bmg_sampler_store": (vp, [vp]),
    "bmg_sampler_chain": (vp, [vp]),
    "bmg_sampler_destroy": (C.c_int, [vp]),
}


ALLGATHER_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_void_p)


class ShardCommStruct(C.Structure):
    """struct bmg_shard_comm (include/bmagwa_b200.h)."""
    _fields_ = [("world", C.c_int), ("rank", C.c_int), ("snp_stride", C.c_int64), ("allgather", ALLGATHER_FN), ("ctx", C.c_void_p)]


class BmgError(RuntimeError):
    pass


def lib():
    global _LIB
    if _LIB is None:
        if not os.path.exists(LIB_PATH):
            raise BmgError("libbmagwa_b200.so is not built (run `python -m bmagwa_b200.build`); there is no CPU fallback")
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            f = getattr(L, name)
            f.restype = res
            f.argtypes = args
        if L.bmg_abi_version() != 1:
            raise BmgError("libbmagwa_b200.so ABI version mismatch")
        _LIB = L
    return _LIB


def check(status):
    if status != 0:
        raise BmgError(lib().bmg_last_error().decode())
