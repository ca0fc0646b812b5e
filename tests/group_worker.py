"""Worker of tests/test_gpu_sharded.py (torch.distributed.run, 2 ranks sharing cuda:0, gloo for the plumbing).

Several chains over one SNP-sharded store (bmg_group_create, BASELINE configs[4]): chain c lives on rank c only, every
rank scans its shard for every chain and stores the dot products into the owning chain's GPU memory.  Checked here:
  mode "two":   2 ranks, 2 chains -- chain r of the group == chain r of a single-GPU run, byte for byte;
  mode "one":   2 ranks, 1 chain  -- rank 1 has no chain and only serves scans (bmg_group_serve);
  mode "lock":  the lockstep single chain (bmg_sampler_create_sharded) over the group's native all-gather."""
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
    work, mode, tau_rng, iters = sys.argv[1], sys.argv[2], sys.argv[3], int(sys.argv[4])
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    torch.cuda.set_device(0)
    n, m_g, n_rao = (int(sys.argv[5]) if len(sys.argv) > 5 else 600), 3000, 100
    miss_rate = float(sys.argv[6]) if len(sys.argv) > 6 else 0.0
    n_chains = {"two": 2, "three": 3}.get(mode, 1)   # three: three ranks, a two-residual pass + a one-residual pass per scan
    if rank == 0:
        ds = synth.write_dataset(work, "syn", n=n, m_g=m_g, m_e=1, seed=5, e_qg=5, var_qg=20, do_n_iter=iters, n_rao=n_rao,
                                 n_rao_burnin=2, n_threads=3, seeds="1234,2345,3456", miss_rate=miss_rate, outbase=os.path.join(work, "single"))
        np.save(os.path.join(work, "y.npy"), ds["y"])
        np.save(os.path.join(work, "E.npy"), ds["E"])
        for c in range(n_chains):   # the same chains on the whole store
            s = api.Sampler(os.path.join(work, "syn.ini"), c, 0, tau_rng=tau_rng)
            s.begin(); s.run(iters); s.end(); s.close()
    dist.barrier()
    ini = os.path.join(work, "syn.ini")
    y = np.load(os.path.join(work, "y.npy"))
    e = np.load(os.path.join(work, "E.npy"))
    store, group = sharded.create_group(dist, n, m_g, 0, y, e, n_chains=world if mode == "lock" else n_chains,
                                        bed_path=os.path.join(work, "syn.bed"))
    if mode == "lock":
        smp = api.Sampler(ini, 0, 0, store=store, comm=group.native_comm(), tau_rng=tau_rng)
        smp.set_option("basename", os.path.join(work, "group%d" % rank))
        smp.begin(); smp.run(iters); smp.end(); smp.close()
    elif group.has_chain:
        smp = api.Sampler(ini, rank, 0, store=store, group=group, tau_rng=tau_rng)
        smp.set_option("basename", os.path.join(work, "group%d" % rank))
        smp.begin()
        smp.run(iters // 2)
        smp.run(iters - iters // 2)
        smp.end()
        assert smp.stats()["scans"] == iters // n_rao
        smp.close()
    else:
        group.serve(iters // n_rao)
    gst = group.stats()
    group.close(); store.close()
    dist.barrier()
    ok = True
    if rank == 0:
        pairs = [("single0", "group0")]
        if mode in ("two", "three"):
            pairs.append(("single1", "group1"))
        if mode == "three":
            pairs.append(("single2", "group2"))
        if mode == "lock":
            pairs.append(("group0", "group1"))
        for a, b in pairs:
            for f in FILES:
                fa, fb = os.path.join(work, a + f), os.path.join(work, b + f)
                if not (os.path.exists(fa) and os.path.exists(fb)):
                    print("missing output", a, b, f); ok = False; continue
                if not filecmp.cmp(fa, fb, shallow=False):
                    xa, xb = np.fromfile(fa, dtype=np.uint8), np.fromfile(fb, dtype=np.uint8)
                    k = min(xa.size, xb.size)
                    d = np.nonzero(xa[:k] != xb[:k])[0]
                    print("%s != %s: %s (sizes %d / %d, first differing byte %s)" % (a, b, f, xa.size, xb.size, d[0] if d.size else "none"))
                    ok = False
        if mode == "two":   # two different seeds: the chains must not be copies of each other
            if filecmp.cmp(os.path.join(work, "group0_loci.dat"), os.path.join(work, "group1_loci.dat"), shallow=False):
                print("chains 0 and 1 are identical"); ok = False
        print("GROUP_OK" if ok else "GROUP_MISMATCH", mode, "rounds", gst["rounds"], "barrier wait %.3f s" % gst["barrier_seconds"])
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
