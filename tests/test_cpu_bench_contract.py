"""Host-side pieces of bench.py that the driver's multi-GPU runs depend on and that no GPU is needed to check."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def test_cpu_blocks_of_the_ranks_are_disjoint_and_prefer_the_gpus_node():
    """bench.plan_cpu_block: every rank gets its own physical cores; ranks whose GPUs hang off one NUMA node share that
    node's cores; unknown or lopsided topologies fall back to contiguous blocks of all allowed cores."""
    import bench
    phys = [[c, c + 32] for c in range(32)]                       # 32 cores x 2 hyperthreads; node 0 = cores 0-15, node 1 = 16-31
    node = [set(range(0, 16)) | set(range(32, 48)), set(range(16, 32)) | set(range(48, 64))]
    local = [node[0]] * 4 + [node[1]] * 4
    blocks = [bench.plan_cpu_block(phys, local, r, 8) for r in range(8)]
    assert all(len(b) == 8 for b in blocks)                       # 4 cores x 2 threads each
    assert len(set(c for b in blocks for c in b)) == 64           # disjoint
    assert all(set(blocks[r]) <= local[r] for r in range(8))      # on the GPU's node
    # one node visible (a VM): the old contiguous split
    one = [set(range(64))] * 8
    assert [bench.plan_cpu_block(phys, one, r, 8) for r in range(8)] == [sorted(c for core in phys[4 * r:4 * r + 4] for c in core) for r in range(8)]
    # topology unknown, or a node with fewer allowed cores than GPUs: contiguous blocks of everything, still disjoint
    for loc in ([set()] * 8, [set(range(0, 3))] * 5 + [node[1]] * 3):
        blocks = [bench.plan_cpu_block(phys, loc, r, 8) for r in range(8)]
        assert len(set(c for b in blocks for c in b)) == sum(len(b) for b in blocks) == 64
    assert bench.plan_cpu_block(phys[:4], [set()] * 8, 0, 8) is None   # fewer cores than ranks: leave the affinity alone
    assert bench.plan_cpu_block(phys, [node[0]], 0, 1) == sorted(node[0])   # one GPU: its node's cores


def test_reference_arm_prints_the_contract_line(tmp_path):
    """`bench.py --impl reference` (the unmodified reference sampler on the host cores, oracle/_ref) on the smallest
    workload: one JSON line with the keys the driver reads, its own cpu_baseline and an e2e that repeats the value."""
    import json
    import subprocess
    from oracle import ref
    if not ref.available():
        import pytest
        pytest.skip("oracle/_ref is not built here")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, BMAGWA_BENCH_DIR=str(tmp_path))
    r = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--workload", "C1s", "--steps", "2",
                        "--warmup", "1", "--n-rao", "100"], capture_output=True, text=True, timeout=300, env=env, cwd=root)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "mcmc_iterations_per_sec" and d["unit"] == "iterations/s"
    assert d["higher_is_better"] is True and d["n_gpus"] == 1 and d["steps"] == 2 and d["warmup"] == 1
    assert d["value"] > 0 and d["gpu_launches"] == 0
    assert d["cpu_baseline"]["kind"] == "reference" and d["cpu_baseline"]["cores"] == 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "iterations/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["config"]["workload"].startswith("C1s")
