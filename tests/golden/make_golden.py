"""Regenerates tests/golden/*.npz from the CPU oracle (run from the repo root: python tests/golden/make_golden.py).

The reference ships no golden vectors and cannot run here, so these fixtures pin the ORACLE against
drift; the oracle itself is pinned by the known-answer tests of tests/test_oracle_kat.py."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from util import golden_case, GOLDEN_DIR  # noqa: E402

for name in ("hybrid_f0", "hybrid_f1", "raytrace", "deferred"):
    out = golden_case(name)
    np.savez_compressed(os.path.join(GOLDEN_DIR, name + ".npz"), **out)
    print(name, {k: v.shape for k, v in out.items()})
