"""Golden chain of the UNMODIFIED reference at the C4 SHAPE (BASELINE configs[3]: n = 50,000 individuals; here with 2,000 SNPs,
which the reference gets through in seconds): 1,500 iterations with the sampler settings of testdata/testdata.ini, three
Rao-Blackwell scans.  The data set is regenerated from the seed by the GPU test (bmagwa_b200/synth.py is deterministic), only
the chain is committed (tests/golden/ref_c4shape_chain.npz).

    python tests/golden/make_c4shape_golden.py"""
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)
from bmagwa_b200 import synth  # noqa: E402
from oracle import ref  # noqa: E402
from make_golden import read_chain  # noqa: E402

SPEC = dict(n=50000, m_g=2000, m_e=2, seed=31, e_qg=20, var_qg=300, do_n_iter=1500, n_rao=500, n_rao_burnin=1000, thin=10,
            n_sample_tau2_and_missing=10, delay_rejection=10, max_move_size=20, use_individual_tau2=1, seeds="1234")


def main():
    assert ref.available(), "build oracle/_ref first (make -C oracle ref)"
    tmp = tempfile.mkdtemp()
    ds = synth.write_dataset(tmp, "c4shape", outbase=os.path.join(tmp, "chain"), **SPEC)
    R = ref.Ref(ds["ini"])
    stats = R.data_stats()
    R.run_chain()
    R.close()
    out = {"data_stats": np.asarray([stats["var_y"], stats["var_x"], stats["mean_x"], stats["yy"]]),
           "payload_md5": np.frombuffer(__import__("hashlib").md5(ds["payload"].tobytes()).digest(), dtype=np.uint8)}
    for k, v in read_chain(os.path.join(tmp, "chain0")).items():
        out[k] = v
    path = os.path.join(HERE, "ref_c4shape_chain.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes; model sizes", out["modelsize"][::15])


if __name__ == "__main__":
    main()
