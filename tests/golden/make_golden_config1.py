"""Golden vector of BASELINE config 1 from the CPU oracle: the reference's default run (1D/MPNP_CO2ER_EDL.py:256-268,
633-796: 50 um mesh, 0.1 M KHCO3, K+, V = -1, dry run = 100 steps of 1e-5 s): Newton count of every step, the OHP
nodal values after steps 10, 50, 100 and the full state after step 100.

    python tests/golden/make_golden_config1.py          # ~1 min
"""
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))

from gmpnp_b200 import meshio, params  # noqa: E402
from oracle import solver  # noqa: E402


def main():
    m = meshio.load_mesh("1D_variable_50um_mesh_5990")
    prm = params.params_1d()
    t = time.time()
    hist, its, _ = solver.march_1d(m.x[:, 0], prm, 100)
    print("oracle: 100 steps,", sum(its), "Newton iterations,", round(time.time() - t, 1), "s")
    np.savez_compressed(os.path.join(HERE, "march_50um_100.npz"), its=np.array(its), last=hist[100],
                        ohp=np.stack([hist[k][0] for k in (10, 50, 100)]))


if __name__ == "__main__":
    main()
