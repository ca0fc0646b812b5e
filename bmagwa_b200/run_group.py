"""`./bmagwa config.ini` over the GPUs of one box: the INI file's `thread.n_threads` chains over ONE store sharded by SNP.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 -m bmagwa_b200.run_group config.ini

The reference runs its chains as threads over one shared `Data` (src/main.cpp:54-108).  Here rank r (one process per GPU)
streams SNPs [r stride, (r+1) stride) of `datafiles.file_g` to its GPU and, for r < n_threads, runs chain r
(`thread.seeds[r]`, output files `<basename><r>_*` as the reference names them); the other ranks only serve the scans
(INTEGRATION.md 10).  n_threads may not exceed the number of ranks.  Every chain writes the bytes its single-GPU run
(`bmagwa_b200/bmagwa config.ini`) writes.  torch.distributed only launches the ranks and carries the start-up handles (gloo;
no NCCL on the data path)."""
import os
import sys
import time


def scan_rounds(do_n_iter, n_rao, n_rao_burnin, flat_proposal_dist):
    """All-SNP scans a chain runs in do_n_iter iterations (src/sampler.cpp:731-745): one per n_rao iterations, none during
    the Rao-Blackwell burn-in when the proposal distribution stays flat."""
    blocks = do_n_iter // n_rao if n_rao > 0 else 0
    return max(0, blocks - n_rao_burnin) if flat_proposal_dist else blocks


def main(argv=None):
    argv = sys.argv[1:] if argv is None else argv
    if len(argv) != 1:
        print("Usage: python -m torch.distributed.run --nproc-per-node N -m bmagwa_b200.run_group INIFILE")
        return 0
    ini = argv[0]
    import torch
    import torch.distributed as dist
    from . import api, sharded
    if not torch.cuda.is_available():
        raise SystemExit("bmagwa_b200.run_group: no CUDA device (there is no CPU path)")
    local_rank = int(os.environ.get("LOCAL_RANK", "0")) % torch.cuda.device_count()   # more ranks than GPUs: they share them
    torch.cuda.set_device(local_rank)
    own_pg = not dist.is_initialized()
    if own_pg:
        dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()

    def opt(section, key, default=""):
        return api.ini_lookup(ini, section, key, default)

    m_g = int(opt("sizes", "m_g", "0"))
    n_chains = int(opt("thread", "n_threads", "1"))
    do_n_iter = int(opt("sampler", "do_n_iter", "0"))
    n_rao = int(opt("sampler", "n_rao", "0"))
    rounds = scan_rounds(do_n_iter, n_rao, int(opt("sampler", "n_rao_burnin", "0")), int(opt("sampler", "flat_proposal_dist", "0")) == 1)
    if n_chains > world:
        raise SystemExit("bmagwa_b200.run_group: thread.n_threads = %d chains need at least %d ranks (one chain per GPU); got %d"
                         % (n_chains, n_chains, world))
    if rank == 0:
        print("-------------------------------------------------------------\n"
              "BMAGWA hot path, B200-native implementation: %d chain(s) over one store sharded over %d GPU(s)\n"
              "-------------------------------------------------------------\n" % (n_chains, world), flush=True)
    t0 = time.perf_counter()
    stride, lo, hi = sharded.shard_range(m_g, world, rank)
    store = api.GenotypeStore.from_ini(ini, lo, hi, local_rank)
    sharded.attach_all_peers(dist, store, world, rank, lo, hi)
    group = sharded.ShardGroup(dist, store, stride, n_chains)
    status = 0
    try:
        if group.has_chain:
            print("Initializing sampler %d (rank %d, SNPs [%d, %d) on GPU %d, ready after %.1f s)" % (rank, rank, lo, hi, local_rank,
                                                                                                  time.perf_counter() - t0), flush=True)
            smp = api.Sampler(ini, rank, local_rank, store=store, group=group)
            smp.begin()
            smp.run(do_n_iter)
            smp.end()
            assert smp.stats()["scans"] == rounds, "scan schedule of the chain differs from the serving ranks'"
            smp.close()
            print("Completed chain %d" % rank, flush=True)
        else:
            group.serve(rounds)
    except Exception as e:   # the other ranks are released by the group's barrier time-out
        print("rank %d: %s" % (rank, e), file=sys.stderr, flush=True)
        status = 1
    group.close()
    store.close()
    if own_pg:
        dist.destroy_process_group()
    return status


if __name__ == "__main__":
    sys.exit(main())
