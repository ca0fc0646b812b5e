import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.getcwd())
from bmagwa_b200 import api, synth
n, m = 5000, 4000
payload, f = synth.make_genotypes(n, m, seed=1)
y, _, _ = synth.make_phenotype(payload, f, n, m, seed=1)
E = np.random.default_rng(0).uniform(size=(n, 2))
st = api.GenotypeStore(payload, n, m, recode_to_minor=True)
st.set_phenotype(y, E)
ch = api.Chain(st)
rs = np.random.default_rng(1)
loci = rs.choice(m, size=25, replace=False).astype(np.int64)
for m_c in (1, 3, 8):
    cand = rs.choice(m, size=m_c, replace=False).astype(np.int64)
    ref = None
    for _ in range(20): ref = ch.column_stats(cand, loci)
    t0 = time.perf_counter()
    for i in range(2000): out = ch.column_stats(cand, loci)
    print("m_c=%d k=25: %.1f us per call (python + ctypes included)" % (m_c, (time.perf_counter() - t0) / 2000 * 1e6), flush=True)
ch.close()
