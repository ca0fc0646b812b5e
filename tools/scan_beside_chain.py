"""Development check (one GPU): scans of a fixed residual, repeated from a second host thread, while a full sampler chain
(moves, its own scans, weight refreshes) runs on the same store.  Every repetition must give the same bits."""
import hashlib
import os
import sys
import tempfile
import threading

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bmagwa_b200 import api, synth  # noqa: E402

n, m = int(sys.argv[1]) if len(sys.argv) > 1 else 20000, int(sys.argv[2]) if len(sys.argv) > 2 else 60000
d = tempfile.mkdtemp()
ds = synth.write_dataset(d, "syn", n=n, m_g=m, m_e=2, seed=3, e_qg=20, var_qg=300, do_n_iter=100000, n_rao=int(os.environ.get("N_RAO", "100")), n_rao_burnin=1000,
                         outbase=os.path.join(d, "chain"))
smp = api.Sampler(ds["ini"], 0, 0, tau_rng="device")
L = smp.L
store = api.GenotypeStore.__new__(api.GenotypeStore)
store.L, store.h, store.n, store.m_g, store.lo, store.hi, store.m, store.m_e = L, L.bmg_sampler_store(smp.h), n, m, 0, m, m, 3
B = api.Chain(store)
B.residual([], [0.1, 0.0, 0.0], [])
ref = B.scan_dots()
h0 = hashlib.md5(ref.tobytes()).hexdigest()
stop = False
bad = [0, 0]


def scans():
    while not stop:
        dd = B.scan_dots()
        bad[1] += 1
        if hashlib.md5(dd.tobytes()).hexdigest() != h0:
            diff = np.nonzero(dd != ref)[0]
            bad[0] += 1
            if bad[0] <= 8:
                print("scan %d: %d dot products differ, SNPs %s" % (bad[1], diff.size, diff[:12]), flush=True)


th = threading.Thread(target=scans)
th.start()
smp.begin()
smp.run(int(os.environ.get("ITERS", "6000")))
stop = True
th.join()
smp.end()
print("scans beside the chain: %d, mismatching: %d" % (bad[1], bad[0]))
B.close()
store.h = None
smp.close()
