"""MCMC inclusion frequencies ("mcmcpos") without a second pass over the chain files.

The reference computes posterior association probabilities offline: `bmagwa_postprocess.py mcmcpos basename nsnps burnin
thin` re-reads every chain's _loci.dat / _modelsize.dat, counts for each SNP the retained samples whose model contains it,
divides by the number of retained samples, writes basenameX_mcmcpos.txt per chain and the plain average over chains to
basename_mcmcpos.txt (bmagwa_postprocess.py:79-123).  The chain files written by this package keep working with that
script unchanged.  Here the sampler keeps the counts while it runs (bmg_sampler_inclusion_counts), and chains running
one per GPU are merged with ONE all-reduce of the per-SNP frequencies:

    counts, ns = sampler.inclusion_counts(m_g)
    p_chain = counts / ns
    p_all = merge_chains(p_chain, dist)          # average over ranks, as the reference averages over chains
    write_mcmcpos(basename + "_mcmcpos.txt", p_all)

mcmcpos_from_files() is the numpy restatement of the reference script (used by the tests as the checker of the
running counts, and usable on chain files of either implementation)."""
import os
import re

import numpy as np


def mcmcpos_from_files(loci_file, nsnps, burnin=0, thin=1):
    """Inclusion frequency per SNP of one chain from its _loci.dat / _modelsize.dat (bmagwa_postprocess.py:94-107)."""
    thin = max(1, int(thin))
    ms = np.fromfile(loci_file.replace("_loci.", "_modelsize."), dtype=np.uint32).astype(np.int64)
    loci = np.fromfile(loci_file, dtype=np.uint32)
    j = np.arange(ms.size)
    keep = (j >= burnin) & ((j - burnin) % thin == 0)
    sample_of_entry = np.repeat(j, ms)
    counts = np.bincount(loci[keep[sample_of_entry]], minlength=nsnps).astype(np.float64)
    nsamples = int(keep.sum())
    return counts / nsamples, nsamples


def mcmcpos(basename, nsnps, burnin=0, thin=1, write=True):
    """The whole `mcmcpos` command: every chain basename<digits>_loci.dat, per-chain files and the chain average."""
    directory, name = os.path.split(basename)
    directory = directory or "."
    files = [x for x in os.listdir(directory) if re.match(re.escape(name) + r"\d+_loci.dat$", x)]
    p = []
    for f in files:
        pi, _ = mcmcpos_from_files(os.path.join(directory, f), nsnps, burnin, thin)
        p.append(pi)
        if write:
            write_mcmcpos(os.path.join(directory, f.replace("_loci.dat", "_mcmcpos.txt")), pi)
    avg = sum(p) / len(p)   # left-to-right sum then one division, as the script's sum(...) / len(p)
    if write:
        write_mcmcpos(os.path.join(directory, name + "_mcmcpos.txt"), avg)
    return avg


def write_mcmcpos(path, p):
    """One value per line in Python's float repr, the script's text format (bmagwa_postprocess.py:109-123)."""
    with open(path, "w") as fh:
        fh.write("\n".join(str(float(v)) for v in p))
        fh.write("\n")


def merge_chains(p_chain, dist=None, device=None):
    """Average of the per-chain frequencies over the ranks of a chain-per-GPU run: one all-reduce (NCCL when the
    process group is NCCL and `device` is a CUDA device; gloo on the CPU).  Without a process group: identity."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return np.asarray(p_chain, dtype=np.float64)
    import torch
    t = torch.as_tensor(np.ascontiguousarray(p_chain, dtype=np.float64))
    if device is not None:
        t = t.to(device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return (t / dist.get_world_size()).cpu().numpy()
