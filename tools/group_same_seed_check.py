"""Development check (torchrun, N >= 2 real GPUs): every rank runs the SAME seed over the sharded store; the chains must stay
byte-identical.  After every Rao-Blackwell period the ranks compare checksums of p_r and the discrete trace so far."""
import hashlib
import os
import sys
import tempfile

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from bmagwa_b200 import _lib, api, sharded  # noqa: E402
import ctypes as C  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    one_gpu = os.environ.get("CHECK_ONE_GPU")   # every rank on cuda:0 (gloo): the same protocol without NVLink
    if one_gpu:
        local = 0
    torch.cuda.set_device(local)
    if one_gpu:
        dist.init_process_group("gloo")
    else:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    n, m, m_e = 50000, int(os.environ.get("CHECK_M", "200000")), 2
    steps, n_rao = int(os.environ.get("CHECK_STEPS", "8")), 500
    stride, lo, hi = sharded.shard_range(m, world, rank)
    payload = bench.device_payload(n, lo, hi, bench.GEN_SEED, torch.device("cuda", local))
    store = api.GenotypeStore(None, n, m, recode_to_minor=True, device=local, snp_lo=lo, snp_hi=hi, payload_device_ptr=payload.data_ptr())
    del payload
    y, E = bench.sharded_phenotype(store, lo, hi, n, m, m_e, dist, world, False)
    tmp = tempfile.mkdtemp(prefix="bmg_check_r%d_" % rank)
    shared = os.path.join(tempfile.gettempdir(), "bmg_check_job")
    if rank == 0:
        bench.write_group_files(shared, n, m, m_e, y, E, n_rao, world)
    dist.barrier()
    ini = bench.group_ini(shared, tmp, rank, n, m, m_e, n_rao, world, True)
    store.set_phenotype(y, E)
    sharded.attach_all_peers(dist, store, world, rank, lo, hi)
    group = sharded.ShardGroup(dist, store, stride, world)
    smp = api.Sampler(ini, rank, local, store=store, group=group, tau_rng="device")
    smp.set_option("basename", os.path.join(tmp, "c"))
    for k, v in (("gram_cache", os.environ.get("CHECK_MEMO")), ("colstats_server", os.environ.get("CHECK_SERVER"))):
        if v is not None:
            smp.set_option(k, v)
    smp.begin()
    L = _lib.lib()
    chain = L.bmg_sampler_chain(smp.h)
    p_r = np.zeros(m)
    for s in range(steps):
        smp.run(n_rao)
        api.check(L.bmg_chain_get_array(chain, 0, p_r.ctypes.data_as(C.POINTER(C.c_double))))
        st = smp.stats()
        digest = hashlib.md5(p_r.tobytes()).hexdigest()
        mine = (digest, int(st["model_size"]), float(st["log_likelihood"]), float(p_r.sum()))
        everyone = [None] * world
        dist.all_gather_object(everyone, mine)
        if rank == 0:
            same = all(e == everyone[0] for e in everyone)
            print("step %d: %s  %s" % (s, "SAME" if same else "DIFFERENT", everyone if not same else everyone[0]), flush=True)
            if not same:
                pr_all = None
        # which SNP range differs?
        t = torch.from_numpy(p_r.copy())
        if not one_gpu:
            t = t.cuda()
        ref = t.clone()
        dist.broadcast(ref, src=world - 1)
        if rank == 0:
            d = (t != ref).nonzero().flatten().cpu().numpy()
            if d.size:
                print("   p_r of rank 0 differs from rank %d's at %d SNPs, first %d last %d (shard boundary %d); max |diff| %.3g"
                      % (world - 1, d.size, d[0], d[-1], stride, float((t - ref).abs().max())), flush=True)
    smp.end(); smp.close()
    dist.barrier()
    group.close(); store.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
