"""Times the genotype-scan kernel in isolation on one GPU (development probe, not the bench).

    python tools/scan_probe.py [n] [m] [reps]

Prints per-variant average kernel time (CUDA events on the chain's stream) and achieved packed-byte
bandwidth against MEASURED_PEAKS.json.  Inputs are generated on the device (random packed codes)."""
import ctypes as C
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bmagwa_b200 import api  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 5000
    m = int(sys.argv[2]) if len(sys.argv) > 2 else 100000
    reps = int(sys.argv[3]) if len(sys.argv) > 3 else 20
    peak = 6466.1
    try:
        peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        pass
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
    bytes_scan = m * B
    out = {}
    for variant in (2, 1, 0):
        ch.set_scan_variant(variant)
        for _ in range(3):
            ch.scan_dots(fetch=False)
        ch.sync()
        times = []
        for _ in range(reps):
            flush.zero_()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            with torch.cuda.stream(stream):
                e0.record(stream)
                ch.scan_dots(fetch=False)
                e1.record(stream)
            ch.sync()
            times.append(e0.elapsed_time(e1))
        t = float(np.median(times))
        out[variant] = t
        print("variant %d: median %.3f ms  min %.3f ms  -> %.1f GB/s packed (%.1f%% of measured %.0f GB/s)" % (
            variant, t, min(times), bytes_scan / t / 1e6, 100 * bytes_scan / t / 1e6 / peak, peak))
    ch.set_scan_variant(1)
    d1 = ch.scan_dots()
    ch.set_scan_variant(0)
    d0 = ch.scan_dots()
    ch.set_scan_variant(2)
    d2 = ch.scan_dots()
    print("variants agree: |v1-v0| %.3g |v2-v0| %.3g max|dot| %.3g" % (float(np.abs(d1 - d0).max()), float(np.abs(d2 - d0).max()), float(np.abs(d1).max())))


if __name__ == "__main__":
    main()
