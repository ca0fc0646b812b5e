"""Worker of tests/test_gpu_sharded.py (launched with torch.distributed.run, 2 ranks sharing cuda:0, gloo).

Every rank runs the SNP-sharded chain; rank 0 also runs the same chain unsharded.  The sharded chain must write
byte-identical output files on every rank and equal the unsharded chain draw for draw: the tensor-core scan is exact
integer arithmetic per SNP, so sharding cannot change a bit of it."""
import filecmp
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bmagwa_b200 import api, sharded, synth  # noqa: E402

FILES = ["_loci.dat", "_modelsize.dat", "_jumpdistance.dat", "_log_likelihood.dat", "_log_prior.dat", "_move_type.dat",
         "_move_size.dat", "_pve.dat", "_alpha.dat", "_sigma2.dat", "_rao.dat"]


def main():
    work, tau_rng, iters = sys.argv[1], sys.argv[2], int(sys.argv[3])
    miss_rate = float(sys.argv[4]) if len(sys.argv) > 4 else 0.0
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    torch.cuda.set_device(0)
    n, m_g = 600, 3000
    if rank == 0:
        ds = synth.write_dataset(work, "syn", n=n, m_g=m_g, m_e=1, seed=5, e_qg=5, var_qg=20, do_n_iter=iters, n_rao=100,
                                 n_rao_burnin=2, miss_rate=miss_rate, outbase=os.path.join(work, "single"))
        np.save(os.path.join(work, "y.npy"), ds["y"])
        np.save(os.path.join(work, "E.npy"), ds["E"])
    dist.barrier()
    ini = os.path.join(work, "syn.ini")
    if rank == 0:
        s = api.Sampler(ini, 0, 0, tau_rng=tau_rng)
        s.begin(); s.run(iters); s.end(); s.close()
    dist.barrier()
    y = np.load(os.path.join(work, "y.npy"))
    e = np.load(os.path.join(work, "E.npy"))
    smp, store, comm = sharded.create_sharded_sampler(dist, ini, n, m_g, os.path.join(work, "syn.bed"), 0, y, e, tau_rng=tau_rng)
    smp.set_option("basename", os.path.join(work, "shard%d" % rank))
    smp.begin(); smp.run(iters); smp.end()
    st = smp.stats()
    smp.close(); store.close()
    dist.barrier()
    ok = True
    if rank == 0:
        for f in FILES:
            a, b, c = (os.path.join(work, p + f) for p in ("single0", "shard0", "shard1"))
            if not (os.path.exists(a) and os.path.exists(b) and os.path.exists(c)):
                print("missing output", f); ok = False; continue
            if not filecmp.cmp(a, b, shallow=False):
                print("sharded != single-GPU:", f); ok = False
            if not filecmp.cmp(b, c, shallow=False):
                print("rank 0 != rank 1:", f); ok = False
        print("SHARDED_OK" if ok else "SHARDED_MISMATCH", "scans", int(st["scans"]), "all-gathers", comm.calls, "bytes", comm.bytes)
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
