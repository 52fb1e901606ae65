"""Per-call cost of a single-field call on device arrays: host submit time (CPU clock, queue never full) and device time per call
(CUDA events around a back-to-back loop, L2-warm).  usage: python scripts/call_overhead.py c1|c3 [direct_max]"""
import sys, time, os
sys.path.insert(0, ".")
import numpy as np, torch
import librmn_b200
from librmn_b200 import configs
name = sys.argv[1]; dmax = int(sys.argv[2]) if len(sys.argv) > 2 else 4
B = librmn_b200.load()
B.set_direct_max_fields(dmax)
c = configs.c1(B, 2048, "linear") if name == "c1" else configs.c3(B, 1, "cubic", extrap="value")
c.defset()
if c.vector:
    u, v = c.source_uv(0)
    di = [torch.from_numpy(u).cuda(), torch.from_numpy(v).cuda()]
    do = [torch.empty(c.dst_shape, device="cuda"), torch.empty(c.dst_shape, device="cuda")]
    call = lambda: B.ezuvint(do[0], do[1], di[0], di[1])
else:
    di = [torch.from_numpy(c.source(0)).cuda()]
    do = [torch.empty(c.dst_shape, device="cuda")]
    call = lambda: B.ezsint(do[0], di[0])
for _ in range(20): call()
torch.cuda.synchronize()
N = 300
t0 = time.perf_counter()
for _ in range(N): call()
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(N): call()
e1.record(); torch.cuda.synchronize()
print("%s direct_max=%d conc=%s: host submit %.1f us/call, wall incl. drain %.1f us/call, device back-to-back %.1f us/call, path %s" % (
    name, dmax, os.environ.get("EZB200_CONCURRENT", "1"), (t1 - t0) / N * 1e6, (t2 - t0) / N * 1e6, e0.elapsed_time(e1) / N * 1e3, B.last_path()))
