"""Regenerates tests/golden/ezref_small.npz from the oracle (oracle/_ref/libezref.so = the reference's own C
sources + oracle/fortran_shim.c).  Run where /root/reference exists:   python tests/golden/make_golden.py
The fixtures pin (a) the oracle against silent drift and (b) the GPU path on boxes where the oracle library
is not present."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from librmn_b200 import EzLib  # noqa: E402
import cases  # noqa: E402

R = EzLib(os.path.join(ROOT, "oracle", "_ref", "libezref.so"))
out = {}
for deg in cases.DEGREES:
    for c in cases.small_cases(R, deg):
        rc, z = cases.run_case(R, c)
        out[c.name + "__rc"] = np.int32(rc)
        out[c.name + "__out"] = z
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "ezref_small.npz"), **out)
print("wrote %d arrays" % len(out))
