"""Writes a Circles initial state (FGBS container) for oracle/_ref/ref_sim: python tools/make_circles_state.py N out.bin"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import fgbs  # noqa: E402

n = int(sys.argv[1])
L = float(np.floor(np.cbrt(float(n)) + 1e-6))
rng = np.random.default_rng(0)
fgbs.write_state(sys.argv[2], {k: rng.uniform(0, L, n).astype(np.float32) for k in ("x", "y", "z")})
print(L)
