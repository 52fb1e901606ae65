"""Short device-resident run of one workload for ncu: python scripts/prof_run.py <workload> <arith> <levels> <steps>"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import librmn_b200, bench
wl, arith, nlev, steps = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
B = librmn_b200.load(); B.set_arith(arith)
case, _, desc = bench.workload(B, wl)
case.defset()
comps = 2 if case.vector else 1
src = case.source_uv(0) if case.vector else (case.source(0),)
d_in = [torch.from_numpy(np.ascontiguousarray(s).ravel()).cuda().repeat(nlev, 1).contiguous() for s in src]
d_out = [torch.empty((nlev, case.n_out), dtype=torch.float32, device="cuda") for _ in range(comps)]
torch.cuda.synchronize()
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for it in range(steps):
    if it == steps - 1: ev0.record()
    rc = B.ezuvint_batch(d_out[0], d_out[1], d_in[0], d_in[1], nlev) if case.vector else B.ezsint_batch(d_out[0], d_in[0], nlev)
ev1.record(); torch.cuda.synchronize()
ms = ev0.elapsed_time(ev1)
print("%s arith=%d levels=%d rc=%d last step %.3f ms -> %.3e pts/s" % (wl, arith, nlev, rc, ms, case.n_out * nlev / ms * 1e3))
