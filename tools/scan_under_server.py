"""Development check (one GPU): is the tensor-core scan bit-reproducible while a persistent column-statistics server of ANOTHER
chain is resident on the same GPU (the situation of a shard group's scan service)?"""
import hashlib
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from bmagwa_b200 import api  # noqa: E402

n, m = int(sys.argv[1]) if len(sys.argv) > 1 else 50000, int(sys.argv[2]) if len(sys.argv) > 2 else 100000
payload = bench.device_payload(n, 0, m, 7, torch.device("cuda", 0))
st = api.GenotypeStore(None, n, m, recode_to_minor=True, payload_device_ptr=payload.data_ptr())
del payload
rs = np.random.default_rng(1)
st.set_phenotype(rs.normal(size=n), rs.uniform(size=(n, 2)))
A, B = api.Chain(st), api.Chain(st)
loci = np.sort(rs.choice(m, size=22, replace=False)).astype(np.int64)
B.residual([], [0.1, 0.0, 0.0], [])
ref = B.scan_dots()
h0 = hashlib.md5(ref.tobytes()).hexdigest()
import threading
bad = 0
stop = False


def hammer():   # chain A's thread: column-statistics requests back to back (ctypes releases the GIL inside the call)
    r2 = np.random.default_rng(2)
    while not stop:
        A.column_stats(r2.choice(m, size=int(r2.integers(1, 6)), replace=False).astype(np.int64), loci)


for with_server in (0, 1, 2):
    th = None
    if with_server == 2:
        th = threading.Thread(target=hammer)
        th.start()
    for rep in range(60):
        if with_server == 1:
            A.column_stats(rs.choice(m, size=2, replace=False).astype(np.int64), loci)   # (re)starts / keeps A's server alive
        d = B.scan_dots()
        if hashlib.md5(d.tobytes()).hexdigest() != h0:
            diff = np.nonzero(d != ref)[0]
            bad += 1
            print("mode %d rep %d: %d of %d dot products differ (SNPs %d .. %d; first: %.17g vs %.17g)" % (with_server, rep, diff.size, m, diff[0], diff[-1], d[diff[0]], ref[diff[0]]))
    if th is not None:
        stop = True
        th.join()
    print("mode %d (0 no server, 1 idle server, 2 server busy with requests from another thread) done, mismatching scans so far: %d" % (with_server, bad), flush=True)
A.close(); B.close(); st.close()
