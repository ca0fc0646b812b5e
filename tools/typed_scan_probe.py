"""Times the scan for several effect types against the additive-only scan on one GPU (development probe).

    python tools/typed_scan_probe.py [n] [m] [reps]

The typed scan (bmg_chain_scan_types) reads the packed store twice -- additive and heterozygote operand of the same
tensor-core kernel -- and derives every type from the two sums; the figure of merit is its time against two passes at the
measured HBM peak.  Inputs are generated on the device (random packed codes, no missing calls)."""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bmagwa_b200 import api  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 5000
    m = int(sys.argv[2]) if len(sys.argv) > 2 else 100000
    reps = int(sys.argv[3]) if len(sys.argv) > 3 else 20
    peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    B = (n + 3) // 4
    g = torch.Generator(device="cuda").manual_seed(1)
    raw = torch.randint(0, 256, (m * B,), dtype=torch.uint8, device="cuda", generator=g)
    raw &= 0b10111011
    y = np.random.default_rng(0).normal(size=n)
    st = api.GenotypeStore(None, n, m, recode_to_minor=True, payload_device_ptr=raw.data_ptr())
    del raw
    st.set_phenotype(y)
    ch = api.Chain(st)
    ch.residual([], [0.0], [])
    stream = torch.cuda.ExternalStream(ch.stream())
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")
    lmp_add, lmp_rem = np.full(5, -3.0), np.full(25, -3.0)
    none = np.zeros(0, dtype=np.int64)

    def timed(fn):
        ts = []
        for _ in range(reps + 3):
            flush.zero_()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            fn()
            e1.record(stream)
            ch.sync()
            ts.append(e0.elapsed_time(e1))
        return float(np.median(ts[3:]))

    t_a = timed(lambda: ch.scan(none, [], [], 0.8, -3.0, -3.0, tau=1.0, tau_mode=0, fetch=False))
    out = {"n": n, "m": m, "additive_scan_ms": t_a}
    for types in ([0, 1], [0, 1, 2, 3], [0, 1, 2, 3, 4]):
        t = timed(lambda: ch.scan_types(types, none, np.zeros(0, dtype=np.int32), np.zeros((0, 2)), np.zeros((0, 2)), 0.8, lmp_add,
                                        lmp_rem, tau_shared=np.ones(4), fetch=False))
        out["types_%s_ms" % "".join(str(v) for v in types)] = t
        print("types %s: %.3f ms (additive-only scan %.3f ms); two passes over %d MB at the measured peak: %.3f ms -> %.0f%% of it"
              % (types, t, t_a, m * B // 1000000, 2 * m * B / peak / 1e6, 100 * (2 * m * B / peak / 1e6) / t))
    print(json.dumps(out))


if __name__ == "__main__":
    main()
