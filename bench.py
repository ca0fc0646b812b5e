#!/usr/bin/env python
"""bench.py -- MCMC iterations/sec of the BMAGWA SNP-inclusion sampler on B200 (BASELINE.json metric),
with the genotype-scan roofline and the reference CPU sampler beside it.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload C2] ...
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...

A STEP is one Rao-Blackwell period of the sampler: n_rao MCMC iterations (moves 0/1/2 with delayed
rejection, tau2/alpha Gibbs every 10, thinning every 10) ending with ONE all-SNP genotype scan and
its epilogue (reference: Sampler::sample loop, src/sampler.cpp:626-834).  value = iterations/sec
summed over the chains of the job.  Sampler/prior settings are those of the reference's bundled
testdata/testdata.ini; data are seeded synthetic genotypes (bmagwa_b200/synth.py).

N = 1: BASELINE.json configs[1] ("C2": n=5,000 x p=100,000, linear, single chain) is the headline line; the
       same JSON line carries sub-records for configs[2] (C3, probit) and configs[3] (C4, n=50,000 x p=1,000,000) on the
       one GPU under "workloads" (value, roofline, clocks, e2e; C4 with a CPU baseline).
N > 1: the split the north star names -- C4 SNP-SHARDED over the N GPUs, N chains over the one sharded store (chain c on
       rank c; every scan served by all ranks through peer memory, bmagwa_b200/csrc/group.cu); scaling "weak" in chains.
       The earlier chain-per-GPU figure on replicated C2 stores is carried as the extra key "replicas_c2".
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: n, m_g, covariates (excluding the constant), description
    "C1s": dict(n=1000, m_g=10000, m_e=2, desc="synthetic stand-in for testdata/testdata.ini (n=1,000 x p=10,000)"),
    "C2": dict(n=5000, m_g=100000, m_e=2, desc="synthetic linear GWAS n=5,000 x p=100,000 SNPs, single chain"),
    "C2x": dict(n=5000, m_g=1000000, m_e=2, desc="synthetic linear GWAS n=5,000 x p=1,000,000 SNPs"),
    "C3": dict(n=10000, m_g=500000, m_e=2, desc="synthetic probit case-control n=10,000 x p=500,000 with latent-variable updates (1.25 GB packed)"),
    "C4": dict(n=50000, m_g=1000000, m_e=2, desc="synthetic linear n=50,000 x p=1,000,000 (12.5 GB packed)"),
    "C5": dict(n=400000, m_g=600000, m_e=2, desc="biobank-scale synthetic n=400,000 x p=600,000 (60 GB packed)"),
    "C4s": dict(n=50000, m_g=200000, m_e=2, desc="synthetic linear n=50,000 x p=200,000 (2.5 GB packed; C4 at one fifth of the SNPs)"),
    "C5s": dict(n=400000, m_g=20000, m_e=2, desc="biobank-scale individuals n=400,000 x p=20,000 (2 GB packed; C5 at one thirtieth of the SNPs)"),
}
GEN_SEED = 20121101
CHAIN_SEEDS = [1234, 2345, 3456, 4567, 5678, 6789, 7890, 8901]
if os.environ.get("BMG_BENCH_SEED_SHIFT"):   # development: which seed the first chain gets
    _k = int(os.environ["BMG_BENCH_SEED_SHIFT"]) % len(CHAIN_SEEDS)
    CHAIN_SEEDS = CHAIN_SEEDS[_k:] + CHAIN_SEEDS[:_k]


def log(*a):
    print(*a, file=sys.stderr, flush=True)


class stdout_to_stderr:
    """The native libraries (ours and the reference) print option warnings on fd 1; keep fd 1 clean so that the
    JSON line is the only thing on stdout."""

    def __enter__(self):
        sys.stdout.flush()
        self.saved = os.dup(1)
        os.dup2(2, 1)
        return self

    def __exit__(self, *exc):
        try:
            import ctypes
            ctypes.CDLL(None).fflush(None)
        except Exception:
            pass
        sys.stdout.flush()
        os.dup2(self.saved, 1)
        os.close(self.saved)
        return False


MISS_RATE = 0.0   # --miss-rate: fraction of genotype calls set missing in the synthetic data (default none, as BASELINE's configs)


def dataset_dir(w, n_threads):
    """The data files do not depend on the number of chains (only the INI written per run does)."""
    base = os.environ.get("BMAGWA_BENCH_DIR", os.path.join(tempfile.gettempdir(), "bmagwa_bench"))
    tag = "" if MISS_RATE == 0.0 else "_miss%g" % MISS_RATE
    return os.path.join(base, "%s_n%d_m%d%s" % (w, WORKLOADS[w]["n"], WORKLOADS[w]["m_g"], tag))


def prepare_dataset(w, n_rao, n_threads, out_dir, do_n_iter):
    """Writes (or reuses) the synthetic data set and an INI file with testdata.ini's settings."""
    from bmagwa_b200 import synth
    spec = WORKLOADS[w]
    d = dataset_dir(w, n_threads)
    os.makedirs(out_dir, exist_ok=True)
    marker = os.path.join(d, "done")
    ini_kw = dict(do_n_iter=do_n_iter, n_rao=n_rao, n_rao_burnin=1000, thin=10, n_sample_tau2_and_missing=10,
                  delay_rejection=10, max_move_size=20, use_individual_tau2=1, e_qg=20, var_qg=300, n_threads=n_threads,
                  seeds=",".join(str(s) for s in CHAIN_SEEDS[:n_threads]), outbase=os.path.join(out_dir, "chain"),
                  verbosity=0)
    if not os.path.exists(marker):
        t0 = time.time()
        synth.write_dataset(d, "syn", n=spec["n"], m_g=spec["m_g"], m_e=spec["m_e"], seed=GEN_SEED, miss_rate=MISS_RATE, **ini_kw)
        open(marker, "w").write("ok")
        log("[bench] generated %s in %.1f s" % (d, time.time() - t0))
    base = os.path.join(d, "syn")
    cfg = dict(base=base, recode=1, n=spec["n"], m_g=spec["m_g"], m_e=spec["m_e"], save_beta=0, types="A")
    cfg.update(ini_kw)
    ini = os.path.join(out_dir, "bench.ini")
    with open(ini, "w") as fh:
        fh.write(synth.INI_TEMPLATE.format(**cfg))
    return ini, spec


def _cpu_list(text):
    out = set()
    for part in text.strip().split(","):
        if part:
            lo, _, hi = part.partition("-")
            out.update(range(int(lo), int(hi or lo) + 1))
    return out


def gpu_local_cpus(device_index):
    """CPUs on the NUMA node the GPU hangs off (sysfs local_cpulist of its PCI function); empty set when unknown."""
    try:
        import torch
        bus = torch.cuda.get_device_properties(device_index).pci_bus_id
        dom = getattr(torch.cuda.get_device_properties(device_index), "pci_domain_id", 0)
        dev = getattr(torch.cuda.get_device_properties(device_index), "pci_device_id", 0)
        path = "/sys/bus/pci/devices/%04x:%02x:%02x.0/local_cpulist" % (dom, bus, dev)
        with open(path) as fh:
            return _cpu_list(fh.read())
    except Exception:
        return set()


def plan_cpu_block(phys, local, local_rank, world):
    """CPUs for one rank.  phys: the allowed physical cores, each a list of its hyperthreads; local[r]: the CPUs on the NUMA
    node of rank r's GPU (empty set = unknown).  Ranks whose GPUs share a node split that node's allowed cores between them;
    when a node has fewer allowed cores than ranks attached to it (or the topology is unknown) every rank gets a contiguous
    block of all allowed cores instead.  Blocks of different ranks never overlap."""
    if len(phys) < world:
        return None
    key = [tuple(sorted(l)) for l in local]
    groups_ok = all(l for l in local)
    if groups_ok:
        for k in set(key):   # every node must hold at least one allowed core per rank attached to it, or nobody uses locality
            members = [r for r in range(world) if key[r] == k]
            near_k = [c for c in phys if set(c) <= set(k)]
            if len(near_k) < len(members):
                groups_ok = False
    if groups_ok:
        peers = [r for r in range(world) if key[r] == key[local_rank]]
        near = [c for c in phys if set(c) <= local[local_rank]]
        per = len(near) // len(peers)
        i = peers.index(local_rank)
        return sorted(c for core in near[i * per:(i + 1) * per] for c in core)
    per = len(phys) // world
    return sorted(c for core in phys[local_rank * per:(local_rank + 1) * per] for c in core)


def pin_rank_to_core(local_rank, world):
    """A block of physical cores (all their hyperthreads) per rank: the chain's host thread spin-waits on device results,
    and two spinning ranks on sibling hyperthreads slow each other down.  Cores on the GPU's own NUMA node are preferred:
    every move is a round trip over PCIe through mapped host memory, and a hop across sockets adds to each leg
    (BMG_BENCH_NO_LOCAL_PIN=1 turns the preference off).  No-op with fewer physical cores than ranks."""
    try:
        allowed = sorted(os.sched_getaffinity(0))
        cores = {}
        for cpu in allowed:
            with open("/sys/devices/system/cpu/cpu%d/topology/thread_siblings_list" % cpu) as fh:
                members = _cpu_list(fh.read())
            cores[min(members)] = sorted(members & set(allowed))
        phys = [cores[k] for k in sorted(cores)]
        if os.environ.get("BMG_BENCH_NO_LOCAL_PIN") is None:
            local = [gpu_local_cpus(r) for r in range(world)]
        else:
            local = [set() for _ in range(world)]
        mine = plan_cpu_block(phys, local, local_rank, world)
        if mine:
            os.sched_setaffinity(0, mine)
        return mine or None
    except Exception:
        return None


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md).  nvidia-smi needs a few
    hundred ms before its first line, more than a short timed region lasts, so it is started before the warm-up steps and
    its time-stamped samples are cut to the timed region afterwards (begin() / end() bracket it)."""
    Q = ("timestamp,index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device_index):
        self.idx = device_index
        self.p = None
        self.t0 = self.t1 = None

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                       "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.p = None

    def begin(self):
        self.t0 = time.time()

    def end(self):
        self.t1 = time.time()

    def stop(self):
        if self.p is None:
            return None
        if self.t1 is None:
            self.end()
        try:
            self.p.terminate()
            try:
                out, _ = self.p.communicate(timeout=5)
            except subprocess.TimeoutExpired:
                self.p.kill()
                out, _ = self.p.communicate()
            return parse_clock_samples(out, self.t0, self.t1)
        except Exception:
            return None


def parse_clock_samples(out, t0, t1):
    """Samples of `nvidia-smi --query-gpu=timestamp,index,clocks.sm,... -lms` inside [t0, t1] (epoch seconds); when the
    region is shorter than the sampling interval and holds none, the sample nearest to it is used and `window` says so."""
    import datetime
    rows = []
    for line in out.splitlines():
        f = [x.strip() for x in line.split(",")]
        if len(f) < 10:
            continue
        try:
            ts = datetime.datetime.strptime(f[0], "%Y/%m/%d %H:%M:%S.%f").timestamp()
            sm, smax = float(f[2]), float(f[3])
        except ValueError:
            continue
        reasons = set(name for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[6:10])
                      if v.lower().startswith("active"))
        rows.append((ts, sm, smax, reasons))
    if not rows:
        return None
    window = "timed region"
    if t0 is not None and t1 is not None:
        inside = [r for r in rows if t0 - 0.02 <= r[0] <= t1 + 0.02]
        if not inside:
            mid = 0.5 * (t0 + t1)
            inside = [min(rows, key=lambda r: abs(r[0] - mid))]
            window = "nearest sample, %.0f ms from the timed region (shorter than the sampling interval)" % (1e3 * abs(inside[0][0] - mid))
        rows = inside
    reasons = set()
    for r in rows:
        reasons |= r[3]
    return {"sm_mhz": float(np.median([r[1] for r in rows])), "sm_max_mhz": float(max(r[2] for r in rows)),
            "reasons": sorted(reasons), "samples": len(rows), "window": window}


def measured_peak():
    try:
        return float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def scan_traffic(workload):
    """dram bytes per scan launch from the committed ncu --set full capture (profiles/), or None."""
    try:
        prof = json.load(open(os.path.join(ROOT, "profiles", "scan_traffic.json")))
        return prof.get(workload)
    except Exception:
        return None


# ------------------------------------------------------------------------------------------------
# reference arm: the unmodified reference sampler on the host cores (oracle/_ref)
# ------------------------------------------------------------------------------------------------
def run_reference_chains(ini, n_chains, n_rao, warmup, steps, scan_probe=None):
    """Returns (seconds of the K timed steps, per-chain wall times).  One host thread per chain over one shared
    Data, as main.cpp does.  The reference's Sampler::sample() cannot be re-entered on a non-empty model (it rebuilds
    its removal distribution as if the model were empty, sampler.cpp:599-606), so "W warm-up steps, then K timed"
    is measured as two deterministic runs of the same seeded chains: T(W + K steps) - T(W steps)."""
    from oracle import ref
    parent = ref.Ref(ini, 0)   # owns the Data; its own sampler is chain 0 of the first set
    opened = [parent]
    parent_used = [False]

    def fresh_set():
        cs = []
        for t in range(n_chains):
            if t == 0 and not parent_used[0]:
                parent_used[0] = True
                cs.append(parent)
            else:
                c = ref.Ref(ini, t, parent=parent)
                opened.append(c)
                cs.append(c)
        return cs

    def phase(chains, iters):
        res = [None] * n_chains

        def work(i):
            chains[i].set_do_n_iter(iters)
            res[i] = chains[i].run_chain()
        th = [threading.Thread(target=work, args=(i,)) for i in range(n_chains)]
        t0 = time.perf_counter()
        for t in th:
            t.start()
        for t in th:
            t.join()
        return time.perf_counter() - t0, res

    t_w = 0.0
    if warmup > 0:
        t_w, _ = phase(fresh_set(), warmup * n_rao)
    t_wk, per = phase(fresh_set(), (warmup + steps) * n_rao)
    if scan_probe is not None:   # the all-SNP scan alone on one thread (p_raoblackwell, src/sampler.cpp:32-261), same Data
        scan_probe.append(parent.scan_time(1))
    for c in reversed(opened):
        c.close()
    return t_wk - t_w, per


def reference_arm(args, rank, world):
    if rank != 0:
        return None
    from oracle import ref
    n_chains = max(1, args.gpus)
    if args.workload is None:
        args.workload = "C2" if args.gpus <= 1 else "C4"
    spec = WORKLOADS[args.workload]
    if args.workload in ("C4", "C5", "C3") and ref.available():
        # our arm's config at N > 1: n_chains chains on the big workload; the reference gets one host thread per chain
        res = sliced_reference(args, args.workload, n_chains, min(args.warmup, 2), min(args.steps, 4))
        return {"metric": "mcmc_iterations_per_sec", "unit": "iterations/s", "impl": "reference", "n_gpus": args.gpus,
                "steps": args.steps, "warmup": args.warmup, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f64", "data": "synthetic", "value": res["value"], "ms_per_step": 1e3 * res["step_seconds_full"],
                "config": {"workload": "%s: %s; %d chain(s), one host thread each (the reference's only parallelism, thread.n_threads)"
                           % (args.workload, spec["desc"], n_chains), "n": spec["n"], "m_g": spec["m_g"], "n_rao": args.n_rao, "chains": n_chains,
                           "step": "%d MCMC iterations per chain incl. one all-SNP scan per chain" % args.n_rao},
                "cpu_baseline": {"value": res["value"], "unit": "iterations/s", "cores": n_chains, "kind": "reference", "sample": res["sample"]},
                "e2e": {"value": res["value"], "unit": "iterations/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
    out = {"metric": "mcmc_iterations_per_sec", "unit": "iterations/s", "impl": "reference", "n_gpus": args.gpus,
           "steps": args.steps, "warmup": args.warmup, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
           "dtype": "f64", "data": "synthetic"}
    if not ref.available():
        out["unavailable"] = "oracle/_ref (the reference compiled against the shims) is not present on this box"
        return out
    with tempfile.TemporaryDirectory() as tmp:
        ini, _ = prepare_dataset(args.workload, args.n_rao, n_chains, tmp, args.n_rao)
        t0 = time.time()
        secs, _ = run_reference_chains(ini, n_chains, args.n_rao, args.warmup, args.steps)
        log("[bench] reference arm total %.1f s" % (time.time() - t0))
    iters = n_chains * args.steps * args.n_rao
    value = iters / secs
    cores = os.cpu_count()
    out.update({
        "value": value, "ms_per_step": 1e3 * secs / args.steps,
        "config": {"workload": "%s: %s; %d chain(s), one host thread each (the reference's only parallelism, thread.n_threads)"
                   % (args.workload, spec["desc"], n_chains), "n": spec["n"], "m_g": spec["m_g"], "n_rao": args.n_rao,
                   "step": "%d MCMC iterations incl. one all-SNP scan" % args.n_rao},
        "cpu_baseline": {"value": value, "unit": "iterations/s", "cores": n_chains, "kind": "reference",
                         "sample": "%d timed iterations per chain after %d warm-up (T(W+K) - T(W) of the same seeded chain), "
                                   "unmodified reference sources built -O3 -march=x86-64-v3 -ffp-contract=off against oracle/shim (OpenBLAS 1 thread per chain), "
                                   "host has %s cores"
                                   % (args.steps * args.n_rao, args.warmup * args.n_rao, cores)},
        "e2e": {"value": value, "unit": "iterations/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    })
    return out


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
def ours_arm(args, rank, local_rank, world, dist=None, workload=None, with_cpu_baseline=True):
    """One chain per rank, every rank holding the whole store (host files -> bmg_sampler_create)."""
    import torch
    from bmagwa_b200 import _lib, api
    import ctypes as C

    workload = workload or args.workload
    dev = local_rank
    torch.cuda.set_device(dev)
    L = _lib.lib()
    spec = WORKLOADS[workload]
    n, m = spec["n"], spec["m_g"]
    tmp = tempfile.mkdtemp(prefix="bmagwa_bench_r%d_" % rank)
    if rank == 0:
        prepare_dataset(workload, args.n_rao, world, tmp, args.n_rao)
    if dist is not None:
        dist.barrier()
    ini, _ = prepare_dataset(workload, args.n_rao, world, tmp, args.n_rao)

    def transfers():
        a, b = C.c_uint64(), C.c_uint64()
        L.bmg_transfer_bytes(C.byref(a), C.byref(b))
        return a.value, b.value

    # ---- device-resident measurement: store already in HBM when the timed region starts
    s = api.Sampler(ini, rank, dev, tau_rng=args.tau_rng)
    s.set_option("basename", os.path.join(tmp, "chain%d" % rank))
    s.begin()
    chain = L.bmg_sampler_chain(s.h)
    stream = torch.cuda.ExternalStream(L.bmg_chain_stream(chain), device=torch.device("cuda", dev))
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")
    L.bmg_chain_scan_kernel_time(chain, 1, None, None, 1)

    def step():
        with torch.cuda.stream(stream):
            flush.zero_()            # evict the 125 MB store from the 126 MB L2 before every step
        s.run(args.n_rao)

    clocks = ClockSampler(dev)
    clocks.start()   # streaming by the time the timed region begins; samples are cut to it afterwards
    for _ in range(args.warmup):
        step()
    torch.cuda.synchronize()
    L.bmg_chain_scan_kernel_time(chain, 1, None, None, 1)
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize()
    clocks.begin()
    launches0 = L.bmg_launch_count()
    h2d0, d2h0 = transfers()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record(stream)
    for _ in range(args.steps):
        step()
    e1.record(stream)
    e1.synchronize()
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    clocks.end()
    ev_ms = e0.elapsed_time(e1)
    clk = clocks.stop()
    launches = L.bmg_launch_count() - launches0
    h2d1, d2h1 = transfers()
    ms_tot, n_l = C.c_double(), C.c_int64()
    L.bmg_chain_scan_kernel_time(chain, 1, C.byref(ms_tot), C.byref(n_l), 0)
    st = s.stats()
    cnt = s.counters()
    s.end()
    s.close()
    elapsed_ms = max(ev_ms, 0.0)
    if dist is not None:
        t = torch.tensor([elapsed_ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        elapsed_ms = float(t.item())
        dist.barrier()

    # ---- end to end through the public API: INI + data files on the host -> store upload + re-coding ->
    #      begin -> K steps -> end (Rao-Blackwell means back on the host)
    torch.cuda.synchronize()
    h2d_a, d2h_a = transfers()
    t0 = time.perf_counter()
    s2 = api.Sampler(ini, rank, dev, tau_rng=args.tau_rng)
    s2.set_option("basename", os.path.join(tmp, "e2e%d" % rank))
    t1 = time.perf_counter()
    s2.begin()
    t2 = time.perf_counter()
    for _ in range(args.steps):
        s2.run(args.n_rao)
    t3 = time.perf_counter()
    s2.end()
    torch.cuda.synchronize()
    e2e_secs = time.perf_counter() - t0
    log("[bench] e2e phases: create %.3f s, begin %.3f s, %d steps %.3f s, end %.3f s"
        % (t1 - t0, t2 - t1, args.steps, t3 - t2, t0 + e2e_secs - t3))
    h2d_b, d2h_b = transfers()
    s2.close()
    if dist is not None:
        t = torch.tensor([e2e_secs], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_secs = float(t.item())

    if rank != 0:
        return None
    iters = world * args.steps * args.n_rao
    value = iters / (elapsed_ms * 1e-3)
    peak, peak_src = measured_peak()
    B = (n + 3) // 4
    bytes_scan = m * B + 8 * m + 8 * n          # packed genotypes once + dot out + residual in
    scan_ms = ms_tot.value / max(1, n_l.value)
    achieved = bytes_scan / (scan_ms * 1e-3) / 1e9 if scan_ms > 0 else 0.0
    out = {
        "metric": "mcmc_iterations_per_sec", "value": value, "unit": "iterations/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": elapsed_ms / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "%s: %s%s" % (workload, spec["desc"], "" if world == 1 else "; %d independent chains, one per GPU, every GPU holding the whole store" % world),
                   "n": n, "m_g": m, "n_rao": args.n_rao, "step": "%d MCMC iterations incl. one all-SNP scan" % args.n_rao,
                   "sampler": "testdata/testdata.ini settings (PMV, thin 10, DR 10, individual tau2)",
                   "tau_rng": args.tau_rng, "miss_rate": args.miss_rate,
                   "per_move": ("column statistics served by the persistent k_colstats_server kernel (host mailbox in mapped pinned "
                                "memory; BMG_COLSTATS_SERVER=0 launches k_column_stats_inline per move instead)")
                   if os.environ.get("BMG_COLSTATS_SERVER", "1") != "0" else "one k_column_stats_inline launch per move",
                   "l2": "256 MiB write on the chain's stream before every step, inside the timed region",
                   "timing": "CUDA events on the chain's stream, max over ranks; wall %.3f ms/step" % (1e3 * wall / args.steps)},
        "e2e": {"value": world * args.steps * args.n_rao / e2e_secs, "unit": "iterations/s",
                "h2d_bytes_per_step": (h2d_b - h2d_a) / args.steps, "d2h_bytes_per_step": (d2h_b - d2h_a) / args.steps,
                "what": "bmg_sampler_create (INI + .fam/.y/.e/.bed on the host -> H2D + device re-coding) + begin + %d steps + end, wall clock"
                        % args.steps},
        "gpu_launches": int(launches),
        "roofline": {"bound": "hbm", "kernel": "k_scan_dots (genotype scan reduction, variant %s)" % os.environ.get("BMG_SCAN_VARIANT", "default"),
                     "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": scan_traffic(workload),
                     "bytes_per_launch": bytes_scan, "avg_launch_ms": scan_ms, "launches_timed": int(n_l.value), "peak_source": peak_src,
                     "traffic_source": "dram__bytes_read.sum + dram__bytes_write.sum of one `ncu --set full` capture of this kernel at this "
                                       "workload, committed under profiles/ (round2_scan_ncu.md, scan_traffic.json): a per-round figure, not "
                                       "measured in this run"},
        "clocks": clk,
        "breakdown": {"move_seconds": st["move_seconds"], "scan_seconds": st["scan_seconds"], "scans": st["scans"], "column_stats_seconds": st["column_stats_seconds"],
                      "delayed_rejection_seconds": cnt["delayed_rejection_seconds"], "delayed_rejection_events": cnt["delayed_rejection_events"],
                      "moves_with_additions": cnt["moves_with_additions"], "served_from_memo": cnt["served_from_memo"],
                      "device_requests": cnt["device_requests"], "server_fallbacks": cnt["server_fallbacks"],
                      "note": "seconds are those of the whole resident chain (warm-up included), counts likewise",
                      "h2d_bytes_per_step_resident": (h2d1 - h2d0) / args.steps, "d2h_bytes_per_step_resident": (d2h1 - d2h0) / args.steps},
    }
    if world == 1 and with_cpu_baseline and not args.no_cpu_baseline:
        out["cpu_baseline"] = cpu_baseline(args, spec, workload)
    return out


# ------------------------------------------------------------------------------------------------
# --sharded: ONE chain over a SNP-sharded store (BASELINE configs[3]); strong scaling
# ------------------------------------------------------------------------------------------------
def device_payload(n, snp_lo, snp_hi, seed, device, block=2048):
    """Seeded PLINK payload of the SNPs [snp_lo, snp_hi) generated on the device: codes 00/10/11 only (no missing
    calls), Binomial(2, f_j) genotypes with f_j ~ U(0.02, 0.95).  Every block of `block` SNPs has its own generator
    seed, so the data do not depend on how the SNP axis is sharded.  Returns a uint8 tensor of
    (snp_hi - snp_lo) * ceil(n/4) bytes."""
    import torch
    B = (n + 3) // 4
    out = torch.empty((snp_hi - snp_lo) * B, dtype=torch.uint8, device=device)
    for blk in range(snp_lo // block, (snp_hi + block - 1) // block):
        g = torch.Generator(device=device).manual_seed(seed * 1000003 + blk)
        f = torch.rand((block, 1), generator=g, device=device) * 0.93 + 0.02
        byte = torch.zeros((block, B), dtype=torch.uint8, device=device)
        for p in range(4):
            a = (torch.rand((block, B), generator=g, device=device) < f).to(torch.uint8)
            b = (torch.rand((block, B), generator=g, device=device) < f).to(torch.uint8)
            x = a + b                            # Binomial(2, f)
            code = x + (x > 0).to(torch.uint8)   # 0 -> 00, 1 -> 10, 2 -> 11
            byte |= code << (2 * p)
        r0, r1 = max(snp_lo, blk * block), min(snp_hi, (blk + 1) * block)
        out[(r0 - snp_lo) * B:(r1 - snp_lo) * B] = byte[r0 - blk * block:r1 - blk * block].reshape(-1)
    return out


def sharded_arm(args, rank, local_rank, world):
    """--lockstep: ONE chain replicated in lockstep over a SNP-sharded store (round 1's mode; strong scaling of the scan only)."""
    import torch
    import torch.distributed as dist
    from bmagwa_b200 import _lib, sharded, synth
    import ctypes as C

    torch.cuda.set_device(local_rank)
    dev = local_rank
    if not dist.is_initialized():
        if world == 1:
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1"); os.environ.setdefault("MASTER_PORT", "29577")
            os.environ.setdefault("RANK", "0"); os.environ.setdefault("WORLD_SIZE", "1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    L = _lib.lib()
    spec = WORKLOADS[args.workload]
    n, m, m_e = spec["n"], spec["m_g"], spec["m_e"]
    stride, lo, hi = sharded.shard_range(m, world, rank)
    t0 = time.perf_counter()
    payload = device_payload(n, lo, hi, GEN_SEED, torch.device("cuda", dev))
    store, stride, lo, hi = sharded.create_shard_store(dist, n, m, dev, payload_device_ptr=payload.data_ptr())
    del payload
    torch.cuda.empty_cache()
    # phenotype: 20 causal SNPs spread over the genome; each owner contributes its columns
    rs = np.random.default_rng(GEN_SEED)
    causal = np.linspace(0, m - 1, 20).astype(np.int64)
    contrib = np.zeros(n)
    swapped = store.counts()[3]
    for j in causal:
        if lo <= j < hi:
            x = store.get_column(int(j), 0)
            if swapped[j - lo]:
                x = 2.0 - x   # effects are positive on the file's allele coding, as in synth.make_phenotype (mixed signs on
                              # the minor-allele coding; same-sign effects make the reference's model grow to > 100 SNPs)
            contrib += 0.14 * (x - x.mean()) / max(x.std(), 1e-9)
    t = torch.from_numpy(contrib).cuda()
    dist.all_reduce(t)
    y = t.cpu().numpy() + rs.normal(size=n) * np.sqrt(0.6)
    if args.probit:
        y = (y > 0).astype(np.float64)   # case-control labels from the liability
    E = rs.uniform(0.0, 1.0, size=(n, m_e))
    tmp = tempfile.mkdtemp(prefix="bmagwa_shard_r%d_" % rank)
    shared = os.path.join(tempfile.gettempdir(), "bmagwa_shard_job")
    if rank == 0:
        os.makedirs(shared, exist_ok=True)
        base = os.path.join(shared, "syn")
        ids = ["per%d per%d" % (i, i) for i in range(n)]
        with open(base + ".fam", "w") as fh:
            fh.write("".join("%s 0 0 1 %.10g\n" % (ident, v) for ident, v in zip(ids, y)))
        with open(base + ".y", "w") as fh:
            fh.write("".join("%s %.17g\n" % (ident, v) for ident, v in zip(ids, y)))
        with open(base + ".e", "w") as fh:
            fh.write("".join("%s %s\n" % (ident, " ".join("%.17g" % v for v in row)) for ident, row in zip(ids, E)))
        open(base + ".bed", "wb").write(bytes([0x6C, 0x1B, 0x01]))   # never read: the shards were generated on the devices
        cfg = dict(base=base, recode=1, n=n, m_g=m, m_e=m_e, save_beta=0, types="A", do_n_iter=args.n_rao, n_rao=args.n_rao, n_rao_burnin=1000,
                   thin=10, n_sample_tau2_and_missing=10, delay_rejection=10, max_move_size=20, use_individual_tau2=1, e_qg=20,
                   var_qg=300, n_threads=1, seeds=str(CHAIN_SEEDS[0]), outbase=os.path.join(shared, "chain"), verbosity=0)
        with open(os.path.join(shared, "bench.ini"), "w") as fh:
            fh.write(synth.INI_TEMPLATE.format(**cfg))
    dist.barrier()
    ini = os.path.join(shared, "bench.ini")
    s, comm = sharded.finish_sharded_sampler(dist, ini, store, stride, lo, hi, dev, y, E, tau_rng=args.tau_rng)
    s.set_option("basename", os.path.join(tmp, "bench%d" % rank))
    if args.probit:
        s.set_option("probit", "1")
    log("[bench] rank %d: shard [%d, %d) ready in %.1f s" % (rank, lo, hi, time.perf_counter() - t0))
    s.begin()
    chain = L.bmg_sampler_chain(s.h)
    stream = torch.cuda.ExternalStream(s.chain_stream())
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")

    def step():
        with torch.cuda.stream(stream):
            flush.zero_()
        s.run(args.n_rao)

    clocks = ClockSampler(dev)
    if rank == 0:
        clocks.start()
    for _ in range(args.warmup):
        step()
    L.bmg_chain_scan_kernel_time(chain, 1, None, None, 1)
    st0 = s.stats()
    launches_a = int(L.bmg_launch_count())
    torch.cuda.synchronize(); dist.barrier()
    clocks.begin()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record(stream)
    for _ in range(args.steps):
        step()
    e1.record(stream)
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    clocks.end()
    dist.barrier()
    clk = clocks.stop() if rank == 0 else None
    t = torch.tensor([e0.elapsed_time(e1)], device="cuda", dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    elapsed_ms = float(t.item())
    launches_b = int(L.bmg_launch_count())
    st1 = s.stats()
    ms_total, n_l = C.c_double(), C.c_int64()
    L.bmg_chain_scan_kernel_time(chain, 0, C.byref(ms_total), C.byref(n_l), 0)
    s.end(); s.close(); store.close()
    try:
        ms_trace = np.fromfile(os.path.join(tmp, "bench%d_modelsize.dat" % rank), dtype=np.uint32)
        ms_trace = [int(v) for v in ms_trace[::max(1, len(ms_trace) // 12)]]
    except Exception:
        ms_trace = None
    peak, peak_src = measured_peak()
    B = (n + 3) // 4
    bytes_scan = (hi - lo) * B + 8 * (hi - lo) + 8 * n      # this rank's shard
    out = None
    if rank == 0:
        avg_ms = ms_total.value / max(1, n_l.value)
        iters = args.steps * args.n_rao
        out = {
            "metric": "mcmc_iterations_per_sec", "value": iters / (elapsed_ms / 1e3), "unit": "iterations/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": elapsed_ms / args.steps, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "%s: %s; ONE chain, store SNP-sharded over %d GPU(s) (%d SNPs per shard)"
                       % (args.workload, spec["desc"], world, stride), "n": n, "m_g": m, "n_rao": args.n_rao,
                       "step": "%d MCMC iterations incl. one all-SNP scan" % args.n_rao, "tau_rng": args.tau_rng,
                       "likelihood": "probit: latent phenotype redrawn on the device every 10 iterations, sigma2 = 1" if args.probit else "linear",
                       "l2": "256 MiB write on the chain's stream before every step, inside the timed region",
                       "collective": "all-gather of the scan's per-SNP p_r (8 B/SNP) through torch.distributed NCCL, once per scan; "
                                     "column statistics read remote shards over CUDA IPC peer mappings",
                       "timing": "CUDA events on the chain's stream, max over ranks; wall %.3f ms/step" % (1e3 * wall / args.steps)},
            "e2e": {"value": iters / wall, "unit": "iterations/s", "h2d_bytes_per_step": None, "d2h_bytes_per_step": None,
                    "what": "wall clock of the K steps through bmg_sampler_run (per-move results and per-scan weights cross PCIe "
                            "inside); the shards are generated on the devices, so there is no store upload to time"},
            "gpu_launches": launches_b - launches_a,
            "roofline": {"bound": "hbm", "kernel": "k_scan_dots_imma on this rank's shard", "achieved": bytes_scan / (avg_ms * 1e-3) / 1e9,
                         "peak": peak, "unit": "GB/s", "frac": bytes_scan / (avg_ms * 1e-3) / 1e9 / peak, "traffic": None,
                         "bytes_per_launch": bytes_scan, "avg_launch_ms": avg_ms, "launches_timed": int(n_l.value), "peak_source": peak_src},
            "clocks": clk,
            "breakdown": {"move_seconds": st1["move_seconds"] - st0["move_seconds"], "scan_seconds": st1["scan_seconds"] - st0["scan_seconds"],
                          "column_stats_seconds": st1["column_stats_seconds"] - st0["column_stats_seconds"],
                          "allgather_calls": comm.calls, "allgather_bytes_received": comm.bytes, "model_size": st1["model_size"], "model_size_trace": ms_trace},
        }
    dist.barrier()
    return out


# ------------------------------------------------------------------------------------------------
# chains over a SNP-sharded store (bmagwa_b200/csrc/group.cu): the north star's multi-GPU split
# ------------------------------------------------------------------------------------------------
def sharded_phenotype(store, lo, hi, n, m, m_e, dist, world, probit):
    """20 causal SNPs spread over the genome, 0.14 s.d. each on the file's allele coding; each owner adds its columns."""
    import torch
    rs = np.random.default_rng(GEN_SEED)
    causal = np.linspace(0, m - 1, 20).astype(np.int64)
    contrib = np.zeros(n)
    swapped = store.counts()[3]
    for j in causal:
        if lo <= j < hi:
            x = store.get_column(int(j), 0)
            if swapped[j - lo]:
                x = 2.0 - x   # effects are positive on the file's allele coding, as in synth.make_phenotype (mixed signs on
                              # the minor-allele coding; same-sign effects make the reference's model grow to > 100 SNPs)
            contrib += 0.14 * (x - x.mean()) / max(x.std(), 1e-9)
    if world > 1:
        t = torch.from_numpy(contrib)
        if dist.get_backend() != "gloo":
            t = t.cuda()
        dist.all_reduce(t)
        contrib = t.cpu().numpy()
    y = contrib + rs.normal(size=n) * np.sqrt(0.6)
    if probit:
        y = (y > 0).astype(np.float64)   # case-control labels from the liability
    E = rs.uniform(0.0, 1.0, size=(n, m_e))
    return y, E


def group_ini(shared, tmp, rank, n, m, m_e, n_rao, n_chains, same_seed):
    """INI of this rank: thread.seeds[rank] is the seed of the chain that lives here.  same_seed: every chain runs seed
    CHAIN_SEEDS[0] (the INI wants unique seeds, so the other positions hold placeholders no rank ever uses)."""
    from bmagwa_b200 import synth
    seeds = list(CHAIN_SEEDS[:n_chains])
    if same_seed:
        seeds = [900000 + i for i in range(n_chains)]
        if rank < n_chains:
            seeds[rank] = CHAIN_SEEDS[0]
    cfg = dict(base=os.path.join(shared, "syn"), recode=1, n=n, m_g=m, m_e=m_e, save_beta=0, types="A", do_n_iter=n_rao, n_rao=n_rao,
               n_rao_burnin=1000, thin=10, n_sample_tau2_and_missing=10, delay_rejection=10, max_move_size=20, use_individual_tau2=1,
               e_qg=20, var_qg=300, n_threads=n_chains, seeds=",".join(str(v) for v in seeds), outbase=os.path.join(tmp, "chain"), verbosity=0)
    ini = os.path.join(tmp, "bench_%s.ini" % ("same" if same_seed else "distinct"))
    with open(ini, "w") as fh:
        fh.write(synth.INI_TEMPLATE.format(**cfg))
    return ini


def write_group_files(shared, n, m, m_e, y, E, n_rao, n_chains):
    from bmagwa_b200 import synth
    os.makedirs(shared, exist_ok=True)
    base = os.path.join(shared, "syn")
    ids = ["per%d per%d" % (i, i) for i in range(n)]
    with open(base + ".fam", "w") as fh:
        fh.write("".join("%s 0 0 1 %.10g\n" % (ident, v) for ident, v in zip(ids, y)))
    with open(base + ".y", "w") as fh:
        fh.write("".join("%s %.17g\n" % (ident, v) for ident, v in zip(ids, y)))
    with open(base + ".e", "w") as fh:
        fh.write("".join("%s %s\n" % (ident, " ".join("%.17g" % v for v in row)) for ident, row in zip(ids, E)))
    open(base + ".bed", "wb").write(bytes([0x6C, 0x1B, 0x01]))   # never read: the shards come from memory
    return base


def group_arm(args, workload, rank, local_rank, world, dist, n_chains, probit=False, with_e2e=True, same_seed=False):
    """n_chains chains over ONE store sharded by SNP over the `world` GPUs (chain c on rank c).
    same_seed: every chain runs the trajectory of the first seed, so that every GPU has the same work whatever N (chains
    with different seeds differ by up to 1.6x in cost per iteration -- model size, adapted move size -- and the slowest one
    sets the max-over-ranks time)."""
    import torch
    from bmagwa_b200 import _lib, api, sharded
    import ctypes as C

    dev = local_rank
    torch.cuda.set_device(dev)
    L = _lib.lib()
    spec = WORKLOADS[workload]
    n, m, m_e = spec["n"], spec["m_g"], spec["m_e"]
    if dist is None:
        dist = sharded.LocalDist()
    stride, lo, hi = sharded.shard_range(m, world, rank)
    t0 = time.perf_counter()
    payload = device_payload(n, lo, hi, GEN_SEED, torch.device("cuda", dev))
    # the shard as a host buffer for the end-to-end run, in pinned memory (pageable if pinning that much is refused)
    host_payload, host_kind = None, None
    if with_e2e:
        try:
            pinned_t = torch.empty(payload.numel(), dtype=torch.uint8, pin_memory=True)
            pinned_t.copy_(payload)
            torch.cuda.synchronize()
            host_payload, host_kind = pinned_t.numpy(), "pinned"
        except Exception as e:
            log("[bench] %s rank %d: pinning %d bytes failed (%r); pageable host buffer" % (workload, rank, payload.numel(), e))
            host_payload, host_kind = payload.cpu().numpy(), "pageable"
    t_gen = time.perf_counter() - t0

    def build(from_host):
        if from_host:
            return api.GenotypeStore(host_payload, n, m, recode_to_minor=True, device=dev, snp_lo=lo, snp_hi=hi)
        return api.GenotypeStore(None, n, m, recode_to_minor=True, device=dev, snp_lo=lo, snp_hi=hi, payload_device_ptr=payload.data_ptr())

    store = build(False)
    del payload
    torch.cuda.empty_cache()
    y, E = sharded_phenotype(store, lo, hi, n, m, m_e, dist, world, probit)
    tmp = tempfile.mkdtemp(prefix="bmagwa_group_r%d_" % rank)
    shared = os.path.join(tempfile.gettempdir(), "bmagwa_group_job_%s_%d" % (workload, world))
    if rank == 0:
        write_group_files(shared, n, m, m_e, y, E, args.n_rao, n_chains)
    dist.barrier()
    ini = group_ini(shared, tmp, rank, n, m, m_e, args.n_rao, n_chains, same_seed)

    def open_group(store, tag):
        ta = time.perf_counter()
        store.set_phenotype(y, E)
        sharded.attach_all_peers(dist, store, world, rank, lo, hi)
        tb = time.perf_counter()
        group = sharded.ShardGroup(dist, store, stride, n_chains)
        tc = time.perf_counter()
        smp = None
        if group.has_chain:
            smp = api.Sampler(ini, rank, dev, store=store, group=group, tau_rng=args.tau_rng)
            smp.set_option("basename", os.path.join(tmp, "%s%d" % (tag, rank)))
            if probit:
                smp.set_option("probit", "1")
        if rank == 0:
            log("[bench] %s rank 0 (%s): phenotype + peers %.3f s, group %.3f s, sampler %.3f s"
                % (workload, tag, tb - ta, tc - tb, time.perf_counter() - tc))
        return group, smp

    group, smp = open_group(store, "bench")
    log("[bench] %s rank %d: shard [%d, %d) generated in %.1f s, ready after %.1f s" % (workload, rank, lo, hi, t_gen, time.perf_counter() - t0))
    if smp is not None:
        smp.begin()
    chain = L.bmg_sampler_chain(smp.h) if smp is not None else group.scan_chain()
    timer_chain = group.scan_chain()   # the scan kernels of every chain over this rank's shard run on the service's stream
    stream = torch.cuda.ExternalStream(L.bmg_chain_stream(chain), device=torch.device("cuda", dev))
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")

    def step(sampler, grp):
        with torch.cuda.stream(stream):
            flush.zero_()
        if sampler is not None:
            sampler.run(args.n_rao)
        else:
            grp.serve(1)

    def transfers():
        a, b = C.c_uint64(), C.c_uint64()
        L.bmg_transfer_bytes(C.byref(a), C.byref(b))
        return a.value, b.value

    def max_over_ranks(v):
        if world == 1:
            return v
        t = torch.tensor([v], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    clocks = ClockSampler(dev)
    if rank == 0:
        clocks.start()
    burn_steps = max(0, args.burnin // args.n_rao)
    t_b = time.perf_counter()
    for _ in range(burn_steps):
        step(smp, group)
    if smp is not None and burn_steps:
        log("[bench] %s rank %d: %d burn-in iterations in %.1f s, model size %d" % (workload, rank, burn_steps * args.n_rao,
                                                                                   time.perf_counter() - t_b, int(smp.stats()["model_size"])))
    for _ in range(args.warmup):
        step(smp, group)
    L.bmg_chain_scan_kernel_time(timer_chain, 1, None, None, 1)
    st0 = smp.stats() if smp is not None else None
    g0 = group.stats()
    launches_a = int(L.bmg_launch_count())
    torch.cuda.synchronize(); dist.barrier()
    clocks.begin()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t1 = time.perf_counter()
    e0.record(stream)
    for _ in range(args.steps):
        step(smp, group)
    e1.record(stream)
    torch.cuda.synchronize()
    wall = time.perf_counter() - t1
    clocks.end()
    dist.barrier()
    clk = clocks.stop() if rank == 0 else None
    elapsed_ms = max_over_ranks(e0.elapsed_time(e1))
    launches_b = int(L.bmg_launch_count())
    st1 = smp.stats() if smp is not None else None
    cnt = smp.counters() if smp is not None else None
    g1 = group.stats()
    ms_total, n_l = C.c_double(), C.c_int64()
    L.bmg_chain_scan_kernel_time(timer_chain, 0, C.byref(ms_total), C.byref(n_l), 0)
    if smp is not None:
        smp.end(); smp.close()
    group.close(); store.close()
    try:
        ms_trace = np.fromfile(os.path.join(tmp, "bench%d_modelsize.dat" % rank), dtype=np.uint32)
        ms_trace = [int(v) for v in ms_trace[::max(1, len(ms_trace) // 12)]]
    except Exception:
        ms_trace = None
    if smp is not None:
        log("[bench] %s rank %d: chain %d model size trace %s, %.1f us per iteration in moves" %
            (workload, rank, rank, ms_trace, 1e6 * (st1["move_seconds"] - st0["move_seconds"]) / (args.steps * args.n_rao)))

    # ---- end to end: the shard as a HOST buffer -> store (H2D + device re-coding) -> group -> sampler -> K steps -> end
    e2e = None
    if with_e2e:
        torch.cuda.synchronize(); dist.barrier()
        h2d_a, d2h_a = transfers()
        t2 = time.perf_counter()
        store2 = build(True)
        t3 = time.perf_counter()
        group2, smp2 = open_group(store2, "e2e")
        if smp2 is not None:
            smp2.begin()
        t4 = time.perf_counter()
        for _ in range(args.steps):
            if smp2 is not None:
                smp2.run(args.n_rao)
            else:
                group2.serve(1)
        if smp2 is not None:
            smp2.end()
        torch.cuda.synchronize()
        e2e_secs = max_over_ranks(time.perf_counter() - t2)
        h2d_b, d2h_b = transfers()
        log("[bench] %s rank %d e2e phases: store from host buffer %.3f s, group + sampler + begin %.3f s, %d steps + end %.3f s"
            % (workload, rank, t3 - t2, t4 - t3, args.steps, time.perf_counter() - t4))
        if smp2 is not None:
            smp2.close()
        group2.close(); store2.close()
        e2e = {"value": n_chains * args.steps * args.n_rao / e2e_secs, "unit": "iterations/s",
               "h2d_bytes_per_step": (h2d_b - h2d_a) / args.steps, "d2h_bytes_per_step": (d2h_b - d2h_a) / args.steps,
               "host_buffer": host_kind,
               "what": "per rank: bmg_store_create from the shard's packed bytes in HOST memory (H2D + device re-coding) + peers + "
                       "bmg_group_create + bmg_sampler_create_grouped + begin + %d steps + end, wall clock, max over ranks; bytes are rank 0's"
                       % args.steps}
    my_ms = e0.elapsed_time(e1)
    per_chain = None
    if world > 1:
        t = torch.zeros(world, device="cuda", dtype=torch.float64)
        t[rank] = (args.steps * args.n_rao / (my_ms * 1e-3)) if smp is not None or rank < n_chains else 0.0
        dist.all_reduce(t)
        per_chain = [float(v) for v in t.cpu().numpy()[:n_chains]]
    dist.barrier()
    if rank != 0:
        return None
    peak, peak_src = measured_peak()
    B = (n + 3) // 4
    avg_ms = ms_total.value / max(1, n_l.value)
    # residuals served per pass over the shard: the chains of a group are scanned two at a time (k_scan_dots_imma2)
    rhs_per_launch = n_chains * (g1["rounds"] - g0["rounds"]) / max(1, n_l.value)
    bytes_scan = int((hi - lo) * B + rhs_per_launch * (8 * (hi - lo) + 8 * n))      # rank 0's shard once + each residual's limbs and results
    iters = n_chains * args.steps * args.n_rao
    out = {
        "metric": "mcmc_iterations_per_sec", "value": iters / (elapsed_ms / 1e3), "unit": "iterations/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": elapsed_ms / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "%s: %s; %d chain(s) over ONE store SNP-sharded over %d GPU(s) (%d SNPs per shard), chain c on rank c"
                   % (workload, spec["desc"], n_chains, world, stride) +
                   ("; every chain runs the trajectory of seed %d, so that every GPU has the same work at every N" % CHAIN_SEEDS[0]
                    if same_seed and n_chains > 1 else ""),
                   "n": n, "m_g": m, "n_rao": args.n_rao, "chains": n_chains,
                   "seeds": "same" if same_seed and n_chains > 1 else "distinct",
                   "n1_point": "`workloads.C4.value` (= `sharded_series_n1.value`) of the `--gpus 1` line: that line's headline is C2",
                   "step": "%d MCMC iterations per chain incl. one all-SNP scan per chain" % args.n_rao, "tau_rng": args.tau_rng,
                   "likelihood": "probit: latent phenotype redrawn on the device every 10 iterations, sigma2 = 1" if probit else "linear",
                   "sampler": "testdata/testdata.ini settings (PMV, thin 10, DR 10, individual tau2)",
                   "burnin": "%d iterations per chain before the warm-up steps, untimed (start-up transient of the chains: models of up to "
                             "~200 SNPs for some seeds, which cost 10x per iteration)" % (burn_steps * args.n_rao),
                   "l2": "256 MiB write on the chain's stream before every step, inside the timed region",
                   "exchange": "per scan round: every rank scans its shard once per chain (limbs pulled over NVLink by peer loads, n x 8 B), every "
                               "chain pulls its dot products from all ranks (peer loads, 8 B per SNP); two host barriers in POSIX shared memory; no "
                               "NCCL and no host callback on the data path; column statistics read remote shards over the peer mappings",
                   "timing": "CUDA events on the chain's stream, max over ranks; wall %.3f ms/step" % (1e3 * wall / args.steps)},
        "gpu_launches": launches_b - launches_a,
        "roofline": {"bound": "hbm", "kernel": "k_scan_dots_imma%s on rank 0's shard (%.2f residuals per pass over the shard)"
                               % ("2" if rhs_per_launch > 1.01 else "", rhs_per_launch), "residuals_per_launch": rhs_per_launch,
                     "achieved": bytes_scan / (avg_ms * 1e-3) / 1e9 if avg_ms > 0 else 0.0, "peak": peak, "unit": "GB/s",
                     "frac": (bytes_scan / (avg_ms * 1e-3) / 1e9 / peak) if avg_ms > 0 else 0.0, "traffic": scan_traffic(workload),
                     "bytes_per_launch": bytes_scan, "avg_launch_ms": avg_ms, "launches_timed": int(n_l.value), "peak_source": peak_src,
                     "note": ("two residuals per pass: the packed bytes are read once for two chains, i.e. %.0f GB/s per residual-pass; at this "
                              "rate the kernel is bound by its consumers' tensor-pipe / issue rate (profiles/round2_scan2_ncu.md), not by HBM; the "
                              "one-residual kernel on the same shard reaches 0.84-0.87" % (rhs_per_launch * bytes_scan / (avg_ms * 1e-3) / 1e9))
                     if rhs_per_launch > 1.01 and avg_ms > 0 else None},
        "clocks": clk,
        "breakdown": {"move_seconds": st1["move_seconds"] - st0["move_seconds"], "scan_seconds": st1["scan_seconds"] - st0["scan_seconds"],
                      "column_stats_seconds": st1["column_stats_seconds"] - st0["column_stats_seconds"],
                      "group_scan_seconds": g1["scan_seconds"] - g0["scan_seconds"], "barrier_wait_seconds": g1["barrier_seconds"] - g0["barrier_seconds"],
                      "scan_rounds": g1["rounds"] - g0["rounds"],
                      "delayed_rejection_seconds": cnt["delayed_rejection_seconds"], "served_from_memo": cnt["served_from_memo"],
                      "moves_with_additions": cnt["moves_with_additions"], "model_size": st1["model_size"], "model_size_trace": ms_trace,
                      "note": "rank 0's chain; seconds between the first and the last timed step"},
    }
    if per_chain is not None:
        out["per_chain_iterations_per_sec"] = per_chain
    if e2e is not None:
        out["e2e"] = e2e
    return out


def sliced_reference(args, workload, n_chains, warmup, steps, m_slice=10000):
    """The unmodified reference sampler on a workload whose start-up passes alone take a quarter of an hour on the host
    (C4: 5e10 genotypes at ~20 ns each): timed on a SLICE of the SNP axis (same n, m_slice SNPs, same
    generator and sampler settings), then the scan's share of a step is scaled to the full SNP count -- the scan is a
    loop over SNPs (src/sampler.cpp:90-259), everything else in a step costs O(n k) or O(m_g) host passes."""
    import torch
    from oracle import ref
    from bmagwa_b200 import synth
    spec = WORKLOADS[workload]
    n, m, m_e = spec["n"], spec["m_g"], spec["m_e"]
    m_slice = min(m_slice, m)
    d = os.path.join(os.environ.get("BMAGWA_BENCH_DIR", os.path.join(tempfile.gettempdir(), "bmagwa_bench")), "%s_slice_n%d_m%d" % (workload, n, m_slice))
    os.makedirs(d, exist_ok=True)
    base = os.path.join(d, "syn")
    t0 = time.time()
    if not os.path.exists(os.path.join(d, "done")):
        dev = torch.device("cuda", 0) if torch.cuda.is_available() else torch.device("cpu")
        payload = device_payload(n, 0, m_slice, GEN_SEED, dev).cpu().numpy()
        with open(base + ".bed", "wb") as fh:
            fh.write(bytes([0x6C, 0x1B, 0x01]))
            fh.write(payload.tobytes())
        B = (n + 3) // 4
        rs = np.random.default_rng(GEN_SEED)
        causal = np.linspace(0, m_slice - 1, 20).astype(np.int64)
        contrib = np.zeros(n)
        for j in causal:
            x = synth.unpack_payload(payload[j * B:(j + 1) * B], n, 1).astype(np.float64).reshape(-1)
            x[x < 0] = 0.0
            contrib += 0.14 * (x - x.mean()) / max(x.std(), 1e-9)
        y = contrib + rs.normal(size=n) * np.sqrt(0.6)
        E = rs.uniform(0.0, 1.0, size=(n, m_e))
        ids = ["per%d per%d" % (i, i) for i in range(n)]
        with open(base + ".fam", "w") as fh:
            fh.write("".join("%s 0 0 1 %.10g\n" % (ident, v) for ident, v in zip(ids, y)))
        with open(base + ".y", "w") as fh:
            fh.write("".join("%s %.17g\n" % (ident, v) for ident, v in zip(ids, y)))
        with open(base + ".e", "w") as fh:
            fh.write("".join("%s %s\n" % (ident, " ".join("%.17g" % v for v in row)) for ident, row in zip(ids, E)))
        del payload
        open(os.path.join(d, "done"), "w").write("ok")
    out_dir = tempfile.mkdtemp(prefix="bmagwa_refslice_")
    cfg = dict(base=base, recode=1, n=n, m_g=m_slice, m_e=m_e, save_beta=0, types="A", do_n_iter=args.n_rao, n_rao=args.n_rao,
               n_rao_burnin=1000, thin=10, n_sample_tau2_and_missing=10, delay_rejection=10, max_move_size=20, use_individual_tau2=1,
               e_qg=20, var_qg=300, n_threads=n_chains, seeds=",".join(str(v) for v in CHAIN_SEEDS[:n_chains]),
               outbase=os.path.join(out_dir, "chain"), verbosity=0)
    ini = os.path.join(out_dir, "bench.ini")
    with open(ini, "w") as fh:
        fh.write(synth.INI_TEMPLATE.format(**cfg))
    t_data = time.time() - t0
    probe = []
    secs, _ = run_reference_chains(ini, n_chains, args.n_rao, warmup, steps, scan_probe=probe)
    t_scan = probe[0]
    step_slice = secs / steps
    step_full = max(step_slice - t_scan, 0.0) + t_scan * (m / float(m_slice))
    log("[bench] sliced reference %s: data %.1f s, %d chain(s): %.3f s per step on the slice of which scan %.3f s -> %.3f s per step at m_g = %d"
        % (workload, t_data, n_chains, step_slice, t_scan, step_full, m))
    return {"value": n_chains * args.n_rao / step_full, "step_seconds_slice": step_slice, "scan_seconds_slice": t_scan,
            "step_seconds_full": step_full, "m_slice": m_slice,
            "sample": "EXTRAPOLATED from a slice: the unmodified reference sampler (oracle/_ref, -O3 -march=x86-64-v3 -ffp-contract=off, "
                      "OpenBLAS 1 thread per chain) run on n = %d x the first %d SNPs of the same generator, %d chain(s) x %d timed iterations "
                      "after %d warm-up; measured %.3f s per %d-iteration step of which %.3f s is the all-SNP scan (timed alone, "
                      "src/sampler.cpp:32-261); the scan's share is scaled by m_g / m_slice = %.0f to %.3f s per step.  The full workload's "
                      "start-up passes alone (5e10 genotypes at ~20 ns, src/data.cpp:324-434) take ~17 min on one core and are not timed; "
                      "host has %s cores" % (n, m_slice, n_chains, steps * args.n_rao, warmup * args.n_rao, step_slice, args.n_rao, t_scan,
                                             m / float(m_slice), step_full, os.cpu_count())}


def cpu_baseline(args, spec, workload=None):
    """The reference sampler (oracle/_ref) on this box's host cores, bounded sample of the same workload."""
    workload = workload or args.workload
    from oracle import ref
    if not ref.available():
        return {"value": None, "unit": "iterations/s", "cores": 0, "kind": "port",
                "sample": "oracle/_ref not present on this box; no sampler-level CPU number (the C oracle restates kernels only)"}
    with tempfile.TemporaryDirectory() as tmp:
        ini, _ = prepare_dataset(workload, args.n_rao, 1, tmp, args.n_rao)
        t0 = time.time()
        secs, _ = run_reference_chains(ini, 1, args.n_rao, 1, 2)   # T(3 steps) - T(1 step)
        log("[bench] cpu_baseline total %.1f s (timed %.2f s)" % (time.time() - t0, secs))
    return {"value": 2 * args.n_rao / secs, "unit": "iterations/s", "cores": 1, "kind": "reference",
            "sample": "%d timed iterations (2 steps) after one warm-up step of one chain of the unmodified reference sampler "
                      "(oracle/_ref: reference sources unmodified, built -O3 -march=x86-64-v3 -ffp-contract=off against oracle/shim, "
                      "OpenBLAS 1 thread); the reference is single-threaded per chain; host has %s cores"
                      % (2 * args.n_rao, os.cpu_count())}


def compact(rec):
    """Sub-record of another workload inside the headline JSON line."""
    keep = ("value", "unit", "ms_per_step", "steps", "warmup", "n_gpus", "config", "e2e", "roofline", "clocks", "gpu_launches", "breakdown",
            "cpu_baseline")
    return {k: rec[k] for k in keep if k in rec}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=None, choices=sorted(WORKLOADS),
                    help="default: C2 (+ C3, C4 sub-records) on one GPU, C4 sharded with one chain per GPU on several")
    ap.add_argument("--n-rao", type=int, default=500, dest="n_rao")
    ap.add_argument("--tau-rng", default="device", choices=["device", "host"], dest="tau_rng")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true", dest="no_e2e", help="sharded workloads: skip the host-buffer end-to-end pass (it keeps a host copy of every shard)")
    ap.add_argument("--no-sub", action="store_true", help="headline workload only (no C3 / C4 / replica sub-records)")
    ap.add_argument("--burnin", type=int, default=20000,
                    help="sharded workloads (C3/C4/C5): MCMC iterations every chain is advanced before the warm-up steps, untimed, so that "
                         "the chains are past their start-up transient (models of up to ~200 SNPs for some seeds) and cost the same")
    ap.add_argument("--chains", type=int, default=0, help="chains over the sharded store (default: one per GPU; e.g. 4 on 8 GPUs for C5)")
    ap.add_argument("--distinct-seeds", action="store_true", dest="distinct_seeds",
                    help="N > 1: the chains of the headline measurement get different seeds (default: all run the first seed's trajectory, "
                         "equal work per GPU; the distinct-seed run is then the sub-record 'distinct_seeds')")
    ap.add_argument("--replicas", action="store_true", help="N > 1: one chain per GPU on replicated stores as the headline (round 1's mode)")
    ap.add_argument("--miss-rate", type=float, default=0.0, dest="miss_rate",
                    help="fraction of genotype calls set missing in the synthetic data (exercises the imputation path; default 0)")
    ap.add_argument("--probit", action="store_true", help="case-control labels + latent-variable updates (sharded workloads; C3 implies it)")
    ap.add_argument("--sharded", "--lockstep", action="store_true", dest="sharded",
                    help="ONE chain replicated in lockstep over a SNP-sharded store (strong scaling of the scan only)")
    args = ap.parse_args()
    global MISS_RATE
    MISS_RATE = args.miss_rate
    if args.warmup < 3 and args.impl == "ours":
        log("[bench] note: fewer than 3 warm-up steps requested")
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world != args.gpus and world == 1 and args.gpus > 1 and args.impl == "ours":
        raise SystemExit("launch with torch.distributed.run --nproc-per-node %d for --gpus %d" % (args.gpus, args.gpus))
    with stdout_to_stderr():
        if args.impl == "reference":
            line = reference_arm(args, rank, world)
        else:
            import torch
            if not torch.cuda.is_available():
                raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback")
            torch.cuda.set_device(local_rank)
            dist = None
            if world > 1:
                import torch.distributed as dist_mod
                dist = dist_mod
                dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
            near = gpu_local_cpus(local_rank)
            log("[bench] rank %d: allowed CPUs %s; CPUs on GPU %d's NUMA node: %s"
                % (rank, sorted(os.sched_getaffinity(0)), local_rank, sorted(near) if near else "unknown"))
            pinned = pin_rank_to_core(local_rank, world)
            if pinned is not None:
                log("[bench] rank %d pinned to CPUs %s" % (rank, pinned))
            if args.sharded:
                args.workload = args.workload or "C4"
                line = sharded_arm(args, rank, local_rank, world)
            elif world == 1:
                wl = args.workload or "C2"
                if wl in ("C3", "C4", "C4s", "C5", "C5s"):
                    line = group_arm(args, wl, rank, local_rank, 1, None, 1, probit=args.probit or wl == "C3", with_e2e=not args.no_e2e)
                else:
                    args.workload = wl
                    line = ours_arm(args, rank, local_rank, world)
                    if not args.no_sub and wl == "C2" and args.miss_rate == 0.0:
                        subs = {}
                        for sub, probit in (("C3", True), ("C4", False)):
                            try:
                                rec = group_arm(args, sub, rank, local_rank, 1, None, 1, probit=probit)
                                if sub == "C4" and not args.no_cpu_baseline:
                                    from oracle import ref
                                    if ref.available():
                                        res = sliced_reference(args, "C4", 1, 1, 2)
                                        rec["cpu_baseline"] = {"value": res["value"], "unit": "iterations/s", "cores": 1, "kind": "reference",
                                                               "sample": res["sample"]}
                                subs[sub] = compact(rec)
                            except Exception as e:   # a sub-record must not cost the headline
                                subs[sub] = {"error": repr(e)}
                                log("[bench] sub-record %s failed: %r" % (sub, e))
                        line["workloads"] = subs
                        if "value" in subs.get("C4", {}):   # --gpus N > 1 runs C4 sharded: its one-GPU point is this sub-record
                            line["sharded_series_n1"] = {"workload": "C4", "value": subs["C4"]["value"], "unit": "iterations/s",
                                                         "note": "the N = 1 point of the `--gpus N > 1` lines (C4, one chain per GPU over one "
                                                                 "SNP-sharded store); the headline of this line is C2, BASELINE's metric config"}
            else:
                wl = args.workload or "C4"
                n_chains = args.chains if args.chains > 0 else world
                if args.replicas:
                    args.workload = args.workload or "C2"
                    line = ours_arm(args, rank, local_rank, world, dist)
                else:
                    line = group_arm(args, wl, rank, local_rank, world, dist, n_chains, probit=args.probit or wl == "C3",
                                     same_seed=not args.distinct_seeds, with_e2e=not args.no_e2e)
                    if not args.no_sub:
                        if not args.distinct_seeds:
                            alt = group_arm(args, wl, rank, local_rank, world, dist, n_chains, probit=args.probit or wl == "C3",
                                            with_e2e=False, same_seed=False)
                            if line is not None and alt is not None:
                                line["distinct_seeds"] = compact(alt)
                                line["distinct_seeds"]["per_chain_iterations_per_sec"] = alt.get("per_chain_iterations_per_sec")
                        rep = ours_arm(args, rank, local_rank, world, dist, workload="C2", with_cpu_baseline=False)
                        if line is not None and rep is not None:
                            line["replicas_c2"] = compact(rep)
            if dist is not None:
                dist.barrier()
                dist.destroy_process_group()
    if line is not None:
        print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()
