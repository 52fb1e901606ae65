"""Tiled-kernel probe: stats + parity vs arith=1 + timings on the bench workloads."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import librmn_b200, bench
B = librmn_b200.load()
B.ezsetopt("verbose", "yes")
for wl, nlev in (("c2", 16), ("c5", 16), ("c1", 1), ("c3", 1)):
    case, _, desc = bench.workload(B, wl)
    case.defset(); B.ezsetopt("verbose", "yes")
    comps = 2 if case.vector else 1
    src = case.source_uv(0) if case.vector else (case.source(0),)
    d_in = [torch.from_numpy(np.ascontiguousarray(s).ravel()).cuda().repeat(nlev, 1).contiguous() for s in src]
    outs = {}
    for arith in (1, 2):
        B.set_arith(arith)
        d_out = [torch.full((nlev, case.n_out), -7.0, dtype=torch.float32, device="cuda") for _ in range(comps)]
        f = (lambda: B.ezuvint_batch(d_out[0], d_out[1], d_in[0], d_in[1], nlev)) if case.vector else (lambda: B.ezsint_batch(d_out[0], d_in[0], nlev))
        rc = f(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5): f()
        e1.record(); torch.cuda.synchronize()
        outs[arith] = [o.clone() for o in d_out]
        print("%s arith=%d rc=%d path=%s %.3f ms/step -> %.3e pts/s" % (wl, arith, rc, B.last_path(), e0.elapsed_time(e1) / 5, case.n_out * nlev / (e0.elapsed_time(e1) / 5) * 1e3), flush=True)
    rng = float(src[0].max() - src[0].min())
    for c in range(comps):
        d = (outs[1][c].double() - outs[2][c].double()).abs()
        print("   comp %d max|a1-a2|/range = %.3e  (unwritten: %d)" % (c, d.max().item() / rng, int((outs[2][c] == -7.0).sum().item())), flush=True)
