"""Round-trip time of bmg_chain_column_stats at a BASELINE shape: python tools/colstats_latency.py n [k] (GPU box)."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from bmagwa_b200 import api  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 50000
k = int(sys.argv[2]) if len(sys.argv) > 2 else 22
m = int(sys.argv[3]) if len(sys.argv) > 3 else 4096
payload = bench.device_payload(n, 0, m, 7, torch.device("cuda", 0))
st = api.GenotypeStore(None, n, m, recode_to_minor=True, payload_device_ptr=payload.data_ptr())
rs = np.random.default_rng(1)
st.set_phenotype(rs.normal(size=n), rs.uniform(size=(n, 2)))
ch = api.Chain(st)
loci = np.sort(rs.choice(m, size=k, replace=False)).astype(np.int64)
if os.environ.get("BMG_PROBE_CLUSTERED"):   # model columns next to each other (what a compact copy of the model's columns would give)
    loci = np.arange(100, 100 + k, dtype=np.int64)
for m_c in (1, 2, 3, 6):
    cands = [rs.choice(m, size=m_c, replace=False).astype(np.int64) for _ in range(400)]
    for c in cands[:50]:
        ch.column_stats(c, loci)
    dt = 0.0
    randloci = os.environ.get("BMG_PROBE_RANDLOCI")
    for c in cands:
        if randloci:   # model columns never seen before: cold in L2 and in the TLB
            loci = np.sort(rs.choice(m, size=k, replace=False)).astype(np.int64)
        t0 = time.perf_counter()
        ch.column_stats(c, loci)
        dt += time.perf_counter() - t0
    dt /= len(cands)
    print("n %d k %d m_c %d: %.2f us per request (ctypes call overhead included)%s" % (n, k, m_c, 1e6 * dt, "  [" + os.environ.get("BMG_NOTE", "") + "]"))
ch.close(); st.close()
