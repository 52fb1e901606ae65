import sys, os, numpy as np
sys.path.insert(0, "tests"); sys.path.insert(0, ".")
import librmn_b200, cases
from librmn_b200 import EzLib
B = librmn_b200.load(); R = EzLib("oracle/_ref/libezref.so")
for arith in (0, 1, 2, 0):
    B.set_arith(arith)
    for cb, cr in zip(cases.small_cases(B, "cubic"), cases.small_cases(R, "cubic")):
        if not cb.name.startswith("C2"): continue
        rb, ob = cases.run_case(B, cb); rr, orr = cases.run_case(R, cr)
        d = np.abs(ob.astype(np.float64) - orr)
        print(cb.name, "arith", arith, "rc", rb, rr, "path", B.last_path(), "max|d|", d.max(), "n>1e-3", int((d > 1e-3).sum()), "b[min,max]", ob.min(), ob.max(), "r[min,max]", orr.min(), orr.max())
        if d.max() > 1e-3:
            print("  rows with error:", np.unique(np.where(d[0] > 1e-3)[0])[:20], " sample b", ob[0, 36, :4], "r", orr[0, 36, :4])
