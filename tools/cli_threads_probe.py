"""How the command line behaves with several chains on ONE GPU (the reference's testdata.ini runs n_threads = 2):
    python tools/cli_threads_probe.py [n_iter] [chain counts, e.g. 1,2,4,8]
Runs bmagwa_b200/bmagwa on the bundled example with 1, 2, 4 and 8 chains and prints wall time and iterations/s
(process start-up, ~2 s, included).  BMG_SERVER_SHARE=0: every chain's server takes one CTA per SM (round-2 behaviour before
the servers of a device shared its SMs)."""
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    from test_gpu_chain import _testdata_ini
    import pathlib
    n_iter = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
    seeds = [1234, 2345, 3456, 4567, 5678, 6789, 7890, 8901]
    counts = [int(v) for v in sys.argv[2].split(",")] if len(sys.argv) > 2 else [1, 2, 4, 8]
    for nt in counts:
        with tempfile.TemporaryDirectory() as d:
            ini = _testdata_ini(pathlib.Path(d), n_iter, n_threads=nt, seeds=",".join(str(s) for s in seeds[:nt]))
            t0 = time.perf_counter()
            try:
                r = subprocess.run([os.path.join(ROOT, "bmagwa_b200", "bmagwa"), ini], capture_output=True, text=True, timeout=240)
                dt = time.perf_counter() - t0
                print("n_threads=%d: rc=%d, %.2f s, %.0f iterations/s aggregate %s" % (nt, r.returncode, dt, nt * n_iter / dt,
                                                                                     r.stderr.strip()[-200:]), flush=True)
            except subprocess.TimeoutExpired:
                print("n_threads=%d: no result after 240 s" % nt, flush=True)


if __name__ == "__main__":
    main()
