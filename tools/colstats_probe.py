"""Times bmg_chain_column_stats round trips (development probe): wall per call and device time between events."""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bmagwa_b200 import api, synth

def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 5000
    m = 4000
    payload, f = synth.make_genotypes(n, m, seed=1)
    y, _, _ = synth.make_phenotype(payload, f, n, m, seed=1)
    E = np.random.default_rng(0).uniform(size=(n, 2))
    st = api.GenotypeStore(payload, n, m, recode_to_minor=True)
    st.set_phenotype(y, E)
    ch = api.Chain(st)
    stream = torch.cuda.ExternalStream(ch.stream())
    rs = np.random.default_rng(1)
    loci = rs.choice(m, size=25, replace=False).astype(np.int64)
    for m_c in (1, 3, 8):
        cand = rs.choice(m, size=m_c, replace=False).astype(np.int64)
        for _ in range(50): ch.column_stats(cand, loci)
        reps = 2000
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
        t0 = time.perf_counter()
        for i in range(reps):
            ev[i][0].record(stream)
            ch.column_stats(cand, loci)
            ev[i][1].record(stream)
        wall = (time.perf_counter() - t0) / reps
        torch.cuda.synchronize()
        dev = np.median([a.elapsed_time(b) for a, b in ev]) * 1e3
        print("n=%d m_c=%d k=25: wall per call %.1f us (incl. ~python/ctypes overhead), device span between events %.1f us" % (n, m_c, wall * 1e6, dev))
        t0 = time.perf_counter()
        for i in range(reps): ch.column_stats(cand, loci)
        print("   without events: %.1f us per call" % ((time.perf_counter() - t0) / reps * 1e6))
    os.environ["BMG_COLSTATS_SLOW"] = "1"
    cand = rs.choice(m, size=3, replace=False).astype(np.int64)
    for _ in range(50): ch.column_stats(cand, loci)
    t0 = time.perf_counter()
    for i in range(1000): ch.column_stats(cand, loci)
    print("slow (memcpy) path m_c=3: %.1f us per call" % ((time.perf_counter() - t0) / 1000 * 1e6))

main()
