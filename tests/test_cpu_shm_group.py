"""The multi-process plumbing of a shard group (several chains over one SNP-sharded store, bmagwa_b200/csrc/group.cu) without a
GPU: the POSIX shared-memory segment and its barrier (csrc/host/shm_group.hpp), exercised by REAL processes."""
import ctypes as C
import multiprocessing as mp
import os
import subprocess
import uuid

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def harness_path(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("shm") / "libshm_harness.so")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-o", out,
                           os.path.join(ROOT, "tests", "harness", "shm_group_harness.cpp"), "-lrt", "-lpthread"])
    return out


def _rank(path, name, rank, world, rounds, q):
    lib = C.CDLL(path)
    lib.harness_shm_rounds.restype = C.c_long
    err = C.create_string_buffer(256)
    rc = lib.harness_shm_rounds(name.encode(), rank, world, C.c_long(rounds), 1, err, 256)
    q.put((rank, int(rc), err.value.decode()))


@pytest.mark.parametrize("world", [2, 4])
def test_barrier_orders_writes_between_processes(harness_path, world):
    """Each of `world` processes writes its round number, meets the others at the barrier and must then see everybody's
    number -- 20,000 rounds; the segment is unlinked as soon as all ranks have mapped it, so nothing is left in /dev/shm."""
    name = "/bmg_test_%s" % uuid.uuid4().hex[:12]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_rank, args=(harness_path, name, r, world, 20000, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=30)
    assert sorted(r for r, _, _ in got) == list(range(world))
    assert all(rc == 0 for _, rc, _ in got), got
    assert not os.path.exists("/dev/shm" + name)


def test_a_missing_peer_times_out_instead_of_hanging(harness_path):
    name = "/bmg_test_%s" % uuid.uuid4().hex[:12]
    env = dict(os.environ, BMG_GROUP_TIMEOUT="1.5")
    code = ("import ctypes as C; lib = C.CDLL(%r); lib.harness_shm_missing_peer.restype = C.c_long; err = C.create_string_buffer(256); "
            "rc = lib.harness_shm_missing_peer(%r, 2, err, 256); print(rc, err.value.decode())" % (harness_path, name.encode()))
    r = subprocess.run(["python", "-c", code], capture_output=True, text=True, timeout=60, env=env)
    assert r.returncode == 0 and r.stdout.startswith("-1") and "timed out" in r.stdout, r.stdout + r.stderr
