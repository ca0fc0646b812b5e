import sys, os, numpy as np, torch
sys.path.insert(0, os.getcwd())
from bmagwa_b200 import api
n, m = int(sys.argv[1]), int(sys.argv[2])
B = (n + 3) // 4
g = torch.Generator(device="cuda").manual_seed(1)
raw = torch.randint(0, 256, (m * B,), dtype=torch.uint8, device="cuda", generator=g)
raw &= 0b10111011
y = np.random.default_rng(0).normal(size=n)
st = api.GenotypeStore(None, n, m, recode_to_minor=True, payload_device_ptr=raw.data_ptr())
st.set_phenotype(y)
ch = api.Chain(st)
ch.residual([], [0.0], [])
ch.set_scan_variant(0); d0 = ch.scan_dots()
ch.set_scan_variant(2)
for rep in range(5):
    d2 = ch.scan_dots()
    bad = np.nonzero(np.abs(d2 - d0) > 1e-9)[0]
    print("rep", rep, "bad", bad.size, bad[:40], (bad[:40] // 16))
    if bad.size:
        print("   d2", d2[bad[:5]], "d0", d0[bad[:5]], "2*sum r", 2*float(np.sum(y - y.mean())), 2*np.sum(np.abs(y-y.mean())))
