"""Repeat small host-array calls many times and compare every result with the first one bit for bit: a call's result must not
depend on what ran before it (transient mismatches = an ordering / lifetime bug between streams, staging buffers or sessions).
usage: python scripts/stress_repeat.py [iterations]"""
import sys, os, hashlib
sys.path.insert(0, "tests"); sys.path.insert(0, ".")
import numpy as np
import librmn_b200, cases
from librmn_b200 import configs
B = librmn_b200.load()
iters = int(sys.argv[1]) if len(sys.argv) > 1 else 100
first = {}
bad = 0
for it in range(iters):
    for degree in ("linear", "cubic"):
        for arith in (2, 0, 2, 1):
            B.set_arith(arith)
            cs = cases.small_cases(B, degree)[:5]
            wind_case = configs.c2(B, 3.0, 2.5, degree); wind_case.vector = True; wind_case.name += "_winds"
            for c in cs + [wind_case]:
                rc, out = cases.run_case(B, c)
                h = hashlib.md5(out.tobytes()).hexdigest()
                key = (c.name, degree, arith)
                if key not in first:
                    first[key] = (rc, h, out.copy())
                elif first[key][:2] != (rc, h):
                    bad += 1
                    d = out != first[key][2]
                    print("MISMATCH it %d %s: rc %d/%d, %d of %d values differ, rows %s" % (it, key, rc, first[key][0], int(d.sum()), d.size,
                          np.unique(np.where(d)[-2])[:8]), flush=True)
            if wind_case.vector:
                u, v = wind_case.source_uv(0)
                so, wo = np.zeros(wind_case.dst_shape, np.float32), np.zeros(wind_case.dst_shape, np.float32)
                wind_case.defset()
                rc = B.ezwdint(so, wo, u, v)
                h = hashlib.md5(so.tobytes() + wo.tobytes()).hexdigest()
                key = ("wdint", degree, arith)
                if key not in first:
                    first[key] = (rc, h, np.stack([so, wo]))
                elif first[key][:2] != (rc, h):
                    bad += 1
                    d = np.stack([so, wo]) != first[key][2]
                    print("MISMATCH it %d %s: %d of %d values differ" % (it, key, int(d.sum()), d.size), flush=True)
B.set_arith(2)
print("iterations %d, keys %d, mismatches %d (EZB200_CONCURRENT=%s EZB200_DIRECT_MAX_FIELDS=%s)" % (iters, len(first), bad,
      os.environ.get("EZB200_CONCURRENT", "1"), os.environ.get("EZB200_DIRECT_MAX_FIELDS", "4")))
