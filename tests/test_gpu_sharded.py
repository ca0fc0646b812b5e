"""one chain over a SNP-sharded store.  Two ranks (sharing the one GPU of the test box; gloo for the
plumbing because NCCL refuses two ranks on one device) must reproduce the single-GPU chain byte for byte."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
@pytest.mark.parametrize("tau_rng,miss_rate", [("host", 0.0), ("device", 0.0), ("host", 0.02), ("device", 0.02)])
def test_sharded_chain_equals_single_gpu_chain(tmp_path, tau_rng, miss_rate):
    """miss_rate > 0: data with missing genotype calls (src/data_model.cpp:78-103): the lockstep ranks draw the same imputed
    values, every rank keeps them for all SNPs over the whole data set's missing-call index."""
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29%03d" % (os.getpid() % 1000), os.path.join(ROOT, "tests", "sharded_worker.py"), str(tmp_path),
           tau_rng, "1200", str(miss_rate)]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "SHARDED_OK" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]


@pytest.mark.gpu
@pytest.mark.parametrize("mode,tau_rng,n,miss_rate", [("two", "host", 600, 0.0), ("two", "device", 600, 0.0), ("one", "device", 600, 0.0),
                                                      ("lock", "host", 600, 0.0), ("two", "device", 9000, 0.0),
                                                      ("two", "host", 600, 0.03), ("one", "device", 600, 0.03), ("lock", "host", 600, 0.03),
                                                      ("three", "device", 600, 0.0), ("three", "host", 9000, 0.02)])
def test_chains_of_a_shard_group_equal_their_single_gpu_runs(tmp_path, mode, tau_rng, n, miss_rate):
    """Several chains over ONE sharded store (bmg_group_create; BASELINE configs[4], reference: chains share one Data,
    src/main.cpp:54-85): chain c on rank c, scans served by every rank through peer memory.  Each chain must write the
    bytes of its single-GPU run; a rank without a chain only serves scans; the lockstep chain also runs over the group's
    native all-gather.  Chains are scanned two at a time (k_scan_dots_imma2); "three" has a pair and a single pass per scan."""
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "3" if mode == "three" else "2",
           "--master-addr", "127.0.0.1",
           "--master-port", "28%03d" % (os.getpid() % 1000), os.path.join(ROOT, "tests", "group_worker.py"), str(tmp_path),
           mode, tau_rng, "1200", str(n), str(miss_rate)]   # n = 9,000: column statistics summed over 9 slices on the device;
    #                                                    miss_rate > 0: missing calls, imputed values of all SNPs on the chain's rank
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "GROUP_OK" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]


def test_shard_range_covers_all_snps():
    from bmagwa_b200.sharded import shard_range
    for m_g in (1, 15, 16, 17, 3000, 100000, 1000003):
        for world in (1, 2, 3, 8):
            got = []
            for r in range(world):
                stride, lo, hi = shard_range(m_g, world, r)
                assert stride % 16 == 0 and stride * world >= m_g and lo == min(m_g, r * stride)
                got += list(range(lo, hi)) if m_g <= 3000 else [lo, hi]
            if m_g <= 3000:
                assert got == list(range(m_g))
            else:
                assert got[0] == 0 and got[-1] == m_g and all(got[2 * i + 1] == got[2 * i + 2] for i in range(world - 1))


@pytest.mark.gpu
def test_run_group_writes_the_files_of_the_command_line(tmp_path):
    """`python -m bmagwa_b200.run_group config.ini` under torchrun (the INI's n_threads chains over one store sharded by SNP,
    shards built by bmg_store_create_from_ini) against `bmagwa_b200/bmagwa config.ini` (the same chains as threads over one
    whole store; reference: src/main.cpp:54-108): every output file byte for byte.  Two ranks, two chains, missing calls."""
    import filecmp
    import shutil
    from bmagwa_b200 import synth
    from bmagwa_b200.run_group import scan_rounds
    assert scan_rounds(1200, 100, 2, False) == 12 and scan_rounds(1200, 100, 2, True) == 10 and scan_rounds(50, 100, 0, False) == 0
    d = str(tmp_path)
    ds = synth.write_dataset(d, "syn", n=600, m_g=3000, m_e=1, seed=5, e_qg=5, var_qg=20, do_n_iter=1200, n_rao=100, n_rao_burnin=2,
                             n_threads=2, seeds="1234,2345", miss_rate=0.02, outbase=os.path.join(d, "cli"))
    cli = os.path.join(ROOT, "bmagwa_b200", "bmagwa")
    r = subprocess.run([cli, ds["ini"]], capture_output=True, text=True, timeout=600, cwd=d)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    ini2 = os.path.join(d, "grp.ini")
    with open(ds["ini"]) as fh:
        text = fh.read()
    assert os.path.join(d, "cli") in text
    with open(ini2, "w") as fh:
        fh.write(text.replace(os.path.join(d, "cli"), os.path.join(d, "grp")))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "27%03d" % (os.getpid() % 1000), "-m", "bmagwa_b200.run_group", ini2]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0 and "Completed chain 1" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]
    names = sorted(f for f in os.listdir(d) if f.startswith("cli") and f.endswith(".dat"))
    assert len(names) >= 20
    for f in names:
        assert filecmp.cmp(os.path.join(d, f), os.path.join(d, "grp" + f[3:]), shallow=False), f
