"""Golden chain of the UNMODIFIED reference on its own bundled example (BASELINE configs[0], /root/reference/testdata):
run in the build container (where /root/reference and oracle/_ref exist), committed as tests/golden/ref_testdata_chain.npz
together with the data files the GPU test feeds to the product's command line (tests/golden/data/testdata.*).

    python tests/golden/make_testdata_golden.py

The INI file is the reference's testdata/testdata.ini verbatim except for the file paths, one chain (seed 1234) and
do_n_iter = 20000 (40 Rao-Blackwell scans, 2,000 thinned samples)."""
import os
import re
import shutil
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)
from oracle import ref  # noqa: E402
from make_golden import read_chain  # noqa: E402

REF_TESTDATA = "/root/reference/testdata"
N_ITER = 20000


def testdata_ini(text, data_dir, out_base, n_iter=N_ITER):
    """The reference's INI text with the data paths, the output prefix, the iteration count and a single chain replaced."""
    def sub(key, value, t):
        return re.sub(r"(?m)^%s\s*=.*$" % re.escape(key), "%s = %s" % (key, value), t)
    for key in ("file_fam", "file_g", "file_e", "file_y"):
        ext = {"file_fam": "fam", "file_g": "bed", "file_e": "e", "file_y": "y"}[key]
        text = sub(key, os.path.join(data_dir, "testdata." + ext), text)
    text = sub("basename", out_base, text)
    text = sub("do_n_iter", str(n_iter), text)
    text = sub("n_threads", "1", text)
    text = sub("seeds", "1234", text)
    text = sub("verbosity", "0", text)
    return text


def main():
    assert ref.available(), "build oracle/_ref first (make -C oracle ref)"
    d = os.path.join(HERE, "data")
    os.makedirs(d, exist_ok=True)
    for ext in ("bed", "fam", "y", "e"):
        shutil.copyfile(os.path.join(REF_TESTDATA, "testdata." + ext), os.path.join(d, "testdata." + ext))
        os.chmod(os.path.join(d, "testdata." + ext), 0o644)
    text = open(os.path.join(REF_TESTDATA, "testdata.ini")).read()
    open(os.path.join(d, "testdata.ini.in"), "w").write(text)   # the reference's settings, paths substituted at test time
    tmp = tempfile.mkdtemp()
    ini = os.path.join(tmp, "t.ini")
    open(ini, "w").write(testdata_ini(text, d, os.path.join(tmp, "chain")))
    R = ref.Ref(ini)
    stats = R.data_stats()
    R.run_chain()
    R.close()
    out = {"data_stats": np.asarray([stats["var_y"], stats["var_x"], stats["mean_x"], stats["yy"]], dtype=np.float64),
           "n_iter": np.asarray([N_ITER])}
    for k, v in read_chain(os.path.join(tmp, "chain0")).items():
        out[k] = v
    path = os.path.join(HERE, "ref_testdata_chain.npz")
    np.savez_compressed(path, **out)
    print("data stats (var y, var x, mean x, ...):", stats)
    print("wrote", path, os.path.getsize(path), "bytes; model size trace", out["modelsize"][::200])


if __name__ == "__main__":
    main()
