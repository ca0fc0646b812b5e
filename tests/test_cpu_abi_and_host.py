"""CPU-only checks: the C-ABI library loads and exports every symbol include/bmagwa_b200.h declares, fails loudly
without a GPU (no CPU fallback), the INI reader / options mirror the reference's validation, the synthetic
PLINK writer round-trips, the reference build (when present) agrees with the oracle, and the N>1 launch path
works over gloo."""
import ctypes as C
import os
import re
import subprocess
import sys

import numpy as np
import pytest

from bmagwa_b200 import synth
from oracle import cpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    from bmagwa_b200 import _lib
    header = open(os.path.join(ROOT, "include", "bmagwa_b200.h")).read()
    declared = set(re.findall(r"\b(bmg_[a-z0-9_]+)\s*\(", header))
    assert len(declared) >= 45
    L = _lib.lib()
    for name in declared:
        assert hasattr(L, name), "libbmagwa_b200.so does not export %s" % name
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    assert L.bmg_abi_version() == 1


def test_no_cpu_fallback_fails_loudly():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from bmagwa_b200 import api
    payload, _ = synth.make_genotypes(20, 5, seed=1)
    with pytest.raises(Exception) as ei:
        api.GenotypeStore(payload, 20, 5)
    assert "no CPU fallback" in str(ei.value) or "CUDA" in str(ei.value)


def test_product_never_imports_the_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "bmagwa_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".hpp")):
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "import oracle" not in text and "from oracle" not in text and "liboracle" not in text, os.path.join(dirpath, f)
                assert "oracle/oracle.c" not in text


def test_bed_writer_round_trip_and_reference_coding():
    rs = np.random.default_rng(0)
    G = rs.integers(-1, 3, size=(37, 11)).astype(np.int8)
    payload = synth.pack_bed_payload(G)
    assert payload.size == 11 * 10
    assert np.array_equal(synth.unpack_payload(payload, 37, 11), G)
    # same decode as the oracle's restatement of data.cpp:36-54
    assert np.array_equal(cpu.decode_matrix(payload, 37, 11, 0), G.astype(np.float64))


def _run_cli(args, cwd=None):
    exe = os.path.join(ROOT, "bmagwa_b200", "bmagwa")
    return subprocess.run([exe] + args, capture_output=True, text=True, cwd=cwd)


def test_cli_usage_and_config_errors(tmp_path):
    r = _run_cli([])
    assert r.returncode == 0 and "Usage:" in r.stdout
    bad = tmp_path / "bad.ini"
    bad.write_text("[sizes]\nn = 10\n")   # m_g missing
    r = _run_cli([str(bad)])
    assert r.returncode != 0 and "Cannot leave option empty: sizes.m_g" in r.stderr
    r = _run_cli([str(tmp_path / "missing.ini")])
    assert r.returncode != 0 and "Cannot load/parse configuration file." in r.stderr


@pytest.mark.parametrize("edit,message", [
    (("n_rao = 500", "n_rao = 505"), "n_rao should be a multiple of thin"),
    (("types = A", "types = Q"), "Unknown model type"),
    (("seeds = 1234", "seeds = 1234,99"), "Number of seeds must equal n_threads"),
    (("e_qg = 5", "e_qg = -1"), "Positive value required for prior.e_qg"),
    (("type = PMV", "type = XYZ"), "unknown sampler type"),
    (("thin = 10", "thin = ten"), "Invalid conversion."),
])
def test_options_validation_messages_match_reference(tmp_path, edit, message):
    ds = synth.write_dataset(str(tmp_path), "syn", n=20, m_g=30, m_e=0, seed=1, e_qg=5, var_qg=20)
    text = open(ds["ini"]).read()
    assert edit[0] in text
    open(ds["ini"], "w").write(text.replace(edit[0], edit[1]))
    r = _run_cli([ds["ini"]])
    assert r.returncode != 0 and message in r.stderr, r.stderr
    from oracle import ref
    if ref.available():   # the reference raises the same text
        with pytest.raises(RuntimeError) as ei:
            ref.Ref(ds["ini"])
        assert message in str(ei.value)


def test_reference_build_matches_oracle_live(ref_lib, tmp_path):
    """oracle.c against the unmodified reference run here (decode, recode, moments, missing, scan)."""
    n, m = 97, 150
    ds = synth.write_dataset(str(tmp_path), "syn", n=n, m_g=m, m_e=1, seed=21, miss_rate=0.02, e_qg=5, var_qg=20,
                             use_individual_tau2=0)
    R = ref_lib.Ref(ds["ini"])
    bed = cpu.read_bed(ds["base"] + ".bed", n, m)
    cpu.recode_minor(bed, n, m)
    for j in range(m):
        assert np.array_equal(cpu.decode_column(bed, n, j, 0), R.get_column(j, 0))
    assert np.array_equal(cpu.moments(bed, n, m), R.moments())
    st = R.data_stats()
    mean, var = cpu.g_var_and_mean(bed, n, m)
    assert (mean, var) == (st["mean_x"], st["var_x"])
    R.model_add(3, 2.5)
    R.model_set_beta_sigma2(np.array([0.1, -0.2, 0.5]), 0.7)
    _, y_hat = R.model_compute_pve()
    p_ref = R.scan(y_hat)
    pp = R.prior_params()
    P = cpu.Prior.make(m, 5, 20)
    off, idx, _ = cpu.missing_index(bed, n, m)
    model_ind = -np.ones(m, dtype=np.int32)
    model_ind[3] = 0
    p = cpu.scan_A(bed, n, m, cpu.moments(bed, n, m), R.y(), y_hat, model_ind, [0.5], [2.5], pp["inv_tau2_alpha2_A"], 0, 0.7,
                   P.log_add([1, 0, 0, 0, 0], 1, 0), P.log_add([0] * 5, 0, 0), miss=(off, idx, np.zeros(idx.size, dtype=np.int8)))
    assert np.abs(p - p_ref).max() < 1e-13
    R.close()


def test_multi_rank_launch_over_gloo(tmp_path):
    """World-size-2 launch on CPU with the gloo backend: the per-rank chain/seed assignment and the max-over-ranks
    timing reduction bench.py uses for N > 1 (no data-path collective: chains are independent)."""
    script = tmp_path / "w.py"
    script.write_text(
        "import os, sys, torch, torch.distributed as dist\n"
        "sys.path.insert(0, %r)\n"
        "import bench\n"
        "dist.init_process_group('gloo')\n"
        "rank, world = dist.get_rank(), dist.get_world_size()\n"
        "seeds = bench.CHAIN_SEEDS[:world]\n"
        "t = torch.tensor([10.0 + rank], dtype=torch.float64)\n"
        "dist.all_reduce(t, op=dist.ReduceOp.MAX)\n"
        "g = [None] * world\n"
        "dist.all_gather_object(g, (rank, seeds[rank]))\n"
        "if rank == 0:\n"
        "    assert t.item() == 10.0 + world - 1\n"
        "    assert sorted(g) == [(0, 1234), (1, 2345)], g\n"
        "    print('gloo ok')\n"
        "dist.destroy_process_group()\n" % ROOT)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
                        "127.0.0.1", "--master-port", "29541", str(script)], capture_output=True, text=True, timeout=240)
    assert r.returncode == 0, r.stderr[-2000:]
    assert "gloo ok" in r.stdout


def test_bench_clock_samples_are_cut_to_the_timed_region():
    """bench.py starts nvidia-smi before the warm-up and keeps the time-stamped samples inside the timed region; a region
    shorter than the sampling interval falls back to the nearest sample and says so."""
    import datetime
    import time
    sys.path.insert(0, ROOT)
    import bench
    now = time.time()

    def ts(t):
        return datetime.datetime.fromtimestamp(t).strftime("%Y/%m/%d %H:%M:%S.%f")[:-3]

    lines = "\n".join("%s, 0, %d, 1965, 400.1, 0x0000000000000000, Not Active, Not Active, Not Active, %s"
                      % (ts(now + 0.05 * i), 1900 + i, "Active" if i == 3 else "Not Active") for i in range(10))
    c = bench.parse_clock_samples(lines, now + 0.1, now + 0.3)
    assert c["samples"] == 5 and c["sm_mhz"] == 1904.0 and c["sm_max_mhz"] == 1965.0
    assert c["reasons"] == ["sw_power_cap"] and c["window"] == "timed region"
    c = bench.parse_clock_samples(lines, now + 1.0, now + 1.01)
    assert c["samples"] == 1 and c["sm_mhz"] == 1909.0 and c["window"].startswith("nearest sample")
    assert bench.parse_clock_samples("garbage\n", now, now + 1) is None
