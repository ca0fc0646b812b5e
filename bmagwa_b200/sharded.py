"""Chains over a SNP-sharded store: host-side plumbing for bmg_sampler_create_sharded (one lockstep chain) and
bmg_group_create / bmg_sampler_create_grouped (several chains, one per rank; SURVEY.md 8e, BASELINE configs[4]).

One process per GPU (torchrun).  Rank r builds the store of SNPs [r*stride, (r+1)*stride), exports its packed shard
over CUDA IPC, attaches every peer's shard (column statistics read remote columns over NVLink), and runs the same
seeded sampler as every other rank; the only collective on the data path is the all-gather of the scan's per-SNP
result (8 B per SNP, once per scan), supplied here through torch.distributed (NCCL: plumbing, not product)."""
import ctypes as C

import numpy as np
import torch

from . import _lib, api


def shard_range(m_g, world, rank, multiple=16):
    """Equal-stride contiguous SNP blocks (stride a multiple of the scan's 16-SNP tile)."""
    stride = -(-m_g // world)
    stride = -(-stride // multiple) * multiple
    lo = min(m_g, rank * stride)
    return stride, lo, min(m_g, lo + stride)


class _DevicePtr:
    def __init__(self, ptr, nbytes):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (int(ptr), False), "version": 2}


class ShardComm:
    """The bmg_shard_comm of this rank.  backend 'nccl': in-place all_gather_into_tensor on the chain's stream;
    'gloo' (tests: several ranks sharing one GPU, which NCCL refuses): staged through host memory."""

    def __init__(self, dist, world, rank, stride, device):
        self.dist, self.world, self.rank, self.stride, self.device = dist, world, rank, stride, device
        self.backend = dist.get_backend()
        self.calls = 0
        self.bytes = 0
        self._cb = _lib.ALLGATHER_FN(self._allgather)
        self.struct = _lib.ShardCommStruct(world, rank, stride, self._cb, None)

    def _allgather(self, ctx, buf, elems, elem_bytes, stream):
        try:
            per = int(elems) * int(elem_bytes)
            full = torch.as_tensor(_DevicePtr(buf, per * self.world), device=torch.device("cuda", self.device))
            mine = full[self.rank * per:(self.rank + 1) * per]
            ext = torch.cuda.ExternalStream(int(stream or 0), device=torch.device("cuda", self.device))
            with torch.cuda.stream(ext):
                if self.backend == "nccl":
                    self.dist.all_gather_into_tensor(full, mine)
                else:
                    ext.synchronize()
                    host = mine.cpu()
                    parts = [torch.empty_like(host) for _ in range(self.world)]
                    self.dist.all_gather(parts, host)
                    full.copy_(torch.cat(parts).to(full.device))
                    ext.synchronize()
            self.calls += 1
            self.bytes += per * (self.world - 1)
            return 0
        except Exception as e:   # never let an exception cross the C boundary
            import sys
            print("[bmagwa_b200.sharded] all-gather failed: %r" % (e,), file=sys.stderr)
            return 1


def attach_all_peers(dist, store, world, rank, lo, hi):
    """Exchange the IPC handles of the packed shards and map every peer's shard into this rank's store."""
    handle, _ = store.export()
    mine = (bytes(handle), int(lo), int(hi))
    everyone = [None] * world
    dist.all_gather_object(everyone, mine)
    for r, (h, plo, phi) in enumerate(everyone):
        if r != rank and phi > plo:
            store.attach_peer(np.frombuffer(h, dtype=np.uint8), plo, phi)


def create_shard_store(dist, n, m_g, device, bed_path=None, payload_device_ptr=None, recode_to_minor=True):
    """This rank's SNP block of the store (phenotype not yet set).  Returns (store, stride, lo, hi)."""
    world, rank = dist.get_world_size(), dist.get_rank()
    stride, lo, hi = shard_range(m_g, world, rank)
    store = api.GenotypeStore(None, n, m_g, recode_to_minor=recode_to_minor, device=device, snp_lo=lo, snp_hi=hi,
                              payload_device_ptr=payload_device_ptr, bed_path=bed_path)
    return store, stride, lo, hi


def finish_sharded_sampler(dist, ini_path, store, stride, lo, hi, device, y, covariates=None, **options):
    """Phenotype, peer shards, communicator and the lockstep sampler.  Returns (sampler, comm)."""
    world, rank = dist.get_world_size(), dist.get_rank()
    store.set_phenotype(y, covariates)
    attach_all_peers(dist, store, world, rank, lo, hi)
    comm = ShardComm(dist, world, rank, stride, device)
    return api.Sampler(ini_path, 0, device, store=store, comm=comm, **options), comm


def create_sharded_sampler(dist, ini_path, n, m_g, bed_path, device, y, covariates=None, recode_to_minor=True,
                           payload_device_ptr=None, **options):
    """Builds shard store + peers + sampler for this rank.  Returns (sampler, store, comm)."""
    store, stride, lo, hi = create_shard_store(dist, n, m_g, device, bed_path=bed_path, payload_device_ptr=payload_device_ptr,
                                               recode_to_minor=recode_to_minor)
    sampler, comm = finish_sharded_sampler(dist, ini_path, store, stride, lo, hi, device, y, covariates, **options)
    return sampler, store, comm


class LocalDist:
    """Stand-in for torch.distributed in a one-rank job (no process group needed)."""

    def get_world_size(self):
        return 1

    def get_rank(self):
        return 0

    def get_backend(self):
        return "local"

    def barrier(self):
        pass

    def all_gather_object(self, out, obj):
        out[0] = obj

    def broadcast_object_list(self, objs, src=0):
        pass


class ShardGroup:
    """bmg_group of this rank: several chains over one SNP-sharded store, chain c on rank c.  Collective constructor.
    torch.distributed only carries the name of the POSIX shared-memory segment the ranks synchronise through; scans
    exchange their data through CUDA-IPC peer memory inside the library.  Collective: creation, every scan (all chains
    of a group run the same schedule; ranks without a chain call serve()), close()."""

    def __init__(self, dist, store, stride, n_chains):
        self.L = _lib.lib()
        self.world, self.rank, self.n_chains, self.stride = dist.get_world_size(), dist.get_rank(), n_chains, stride
        name = [None]
        if self.rank == 0:
            import os
            import uuid
            name[0] = "/bmg_%d_%s" % (os.getpid(), uuid.uuid4().hex[:12])
        if self.world > 1:
            dist.broadcast_object_list(name, src=0)
        self.name = name[0]
        h = api.vp()
        api.check(self.L.bmg_group_create(store.h, self.world, self.rank, n_chains, stride, self.name.encode(), C.byref(h)))
        self.h = h
        self._store = store

    @property
    def has_chain(self):
        return self.rank < self.n_chains

    def serve(self, n_rounds):
        api.check(self.L.bmg_group_serve(self.h, n_rounds))

    def scan_chain(self):
        return self.L.bmg_group_scan_chain(self.h)

    def stats(self):
        out = np.zeros(4)
        api.check(self.L.bmg_group_stats(self.h, out.ctypes.data_as(api.f64p)))
        return {"rounds": int(out[0]), "scan_seconds": float(out[1]), "scans": int(out[2]), "barrier_seconds": float(out[3])}

    def native_comm(self):
        """bmg_shard_comm for the lockstep single chain over this group's all-gather (no host callback)."""
        class _Native:
            pass
        nc = _Native()
        nc.struct = _lib.ShardCommStruct(self.world, self.rank, self.stride, _lib.ALLGATHER_FN(), C.cast(self.h, C.c_void_p))
        nc.calls = nc.bytes = 0
        return nc

    def close(self):
        if getattr(self, "h", None):
            self.L.bmg_group_destroy(self.h)
            self.h = None


def create_group(dist, n, m_g, device, y, covariates=None, n_chains=None, bed_path=None, payload_device_ptr=None,
                 recode_to_minor=True):
    """Shard store of this rank (phenotype set, peers attached) and the shard group.  Returns (store, group)."""
    world, rank = dist.get_world_size(), dist.get_rank()
    store, stride, lo, hi = create_shard_store(dist, n, m_g, device, bed_path=bed_path, payload_device_ptr=payload_device_ptr,
                                               recode_to_minor=recode_to_minor)
    store.set_phenotype(y, covariates)
    attach_all_peers(dist, store, world, rank, lo, hi)
    group = ShardGroup(dist, store, stride, world if n_chains is None else n_chains)
    return store, group
