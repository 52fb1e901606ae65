"""Threaded against serial calls on shared grids, with the number and location of differing values (the diagnostic behind\ntests/test_threads_gpu.py).  usage: python scripts/threads_vs_serial.py"""
import sys, os, threading
sys.path.insert(0, "tests"); sys.path.insert(0, ".")
import numpy as np
import librmn_b200
from librmn_b200 import configs
import test_threads_gpu as tt
B = librmn_b200.load()
for rep in range(4):
    cs = tt._cases(B)
    jobs = [(c, i + rep, 3) for i, c in enumerate(cs)]
    got = tt._threaded(B, jobs)
    serial = [tt._run(B, *j) for j in jobs]
    again = tt._threaded(B, jobs)
    for j, (rc_s, o_s), (rc_t, o_t), (rc_a, o_a) in zip(jobs, serial, got, again):
        d1 = o_s != o_t; d2 = o_s != o_a
        msg = "rep %d %-28s first-threaded vs serial: %d differ; second-threaded vs serial: %d differ" % (rep, j[0].name, int(d1.sum()), int(d2.sum()))
        if d1.any():
            idx = np.argwhere(d1)
            msg += " | comps %s fields %s rows %d..%d cols %d..%d max|d| %.3g" % (np.unique(idx[:, 0]) if o_s.ndim == 4 else "-", np.unique(idx[:, -3]), idx[:, -2].min(), idx[:, -2].max(),
                                                                         idx[:, -1].min(), idx[:, -1].max(), float(np.abs(o_s.astype(np.float64) - o_t)[d1].max()))
        print(msg, flush=True)
print("env", {k: v for k, v in os.environ.items() if k.startswith("EZB200")})
