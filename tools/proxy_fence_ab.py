"""A/B of the proxy fence in the scan kernel's stage ring (scan_imma.cu): repeated launches of the same scan must give
the same bits.  Run twice, with and without BMG_IMMA_NO_PROXY_FENCE=1:
    python tools/proxy_fence_ab.py [launches]
Prints, per shape, how many launches differed from the first one and which rows of the 16-SNP tiles were hit."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    import torch
    from bmagwa_b200 import api
    launches = int(sys.argv[1]) if len(sys.argv) > 1 else 200
    mode = "WITHOUT the fence" if os.environ.get("BMG_IMMA_NO_PROXY_FENCE") else "with the fence"
    for n, m in ((5120, 20000), (5000, 100000), (3000, 50000), (50000, 20000)):
        B = (n + 3) // 4
        g = torch.Generator(device="cuda").manual_seed(n + m)
        raw = torch.randint(0, 256, (m * B,), dtype=torch.uint8, device="cuda", generator=g)
        raw &= 0b10111011
        y = np.random.default_rng(n).normal(size=n)
        st = api.GenotypeStore(None, n, m, recode_to_minor=True, payload_device_ptr=raw.data_ptr())
        del raw
        st.set_phenotype(y)
        ch = api.Chain(st)
        ch.residual([], [0.0], [])
        ch.set_scan_variant(0)
        ref = ch.scan_dots()
        ch.set_scan_variant(2)
        bad, rows = 0, np.zeros(16, dtype=np.int64)
        worst = 0.0
        for i in range(launches):
            if i % 2 == 0:
                torch.cuda.synchronize()   # every other launch starts on an idle GPU
            d = ch.scan_dots()
            diff = np.nonzero(np.abs(d - ref) > 1e-9 * np.abs(ref).max())[0]
            if diff.size:
                bad += 1
                np.add.at(rows, diff % 16, 1)
                worst = max(worst, float(np.abs(d - ref).max()))
        print("%s: n=%d m=%d: %d of %d launches wrong; SNPs hit by row of their tile %s; largest error %.3g (|dot| up to %.3g)"
              % (mode, n, m, bad, launches, rows.tolist(), worst, float(np.abs(ref).max())), flush=True)
        ch.close(); st.close()


if __name__ == "__main__":
    main()
