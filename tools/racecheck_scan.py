"""compute-sanitizer --tool racecheck python tools/racecheck_scan.py : shared-memory hazards of the tensor-core scan."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bmagwa_b200 import api, synth
n, m = int(sys.argv[1]) if len(sys.argv) > 1 else 5000, int(sys.argv[2]) if len(sys.argv) > 2 else 3000
payload, f = synth.make_genotypes(n, m, seed=3)
y, _, _ = synth.make_phenotype(payload, f, n, m, seed=3)
st = api.GenotypeStore(payload, n, m, recode_to_minor=True, device=0)
st.set_phenotype(y)
ch = api.Chain(st)
ch.residual([], [0.1], [])
d = ch.scan_dots()
print("scan done", float(d.sum()))
ch.close(); st.close()
