"""Small run of every interpolation kernel family for compute-sanitizer (memcheck / racecheck / initcheck): row-separable + window +
list kernels with tail stages, the few-field tile variant, the direct kernel (scalar, fused (u, v) pair, constant extrapolation,
zone list, row band), the constant-store split of the tile list, arith 0 / 1 per-point kernels on a Yin-Yang source."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import librmn_b200
from librmn_b200 import configs
B = librmn_b200.load()
B.set_arith(2)
for dmax in (0, 4):
    B.set_direct_max_fields(dmax)
    for mk, nf in ((lambda: configs.c2(B, 3.0, 2.0, "cubic"), 6), (lambda: configs.c5(B, 3.0, 2.0, "cubic"), 5), (lambda: configs.c1(B, 96, "linear"), 3),
                   (lambda: configs.c5(B, 3.0, 2.0, "cubic"), 1), (lambda: configs.c2(B, 3.0, 2.0, "linear"), 2)):
        c = mk(); c.defset()
        zin = torch.from_numpy(np.stack([c.source(k) for k in range(nf)])).cuda().contiguous()
        z = torch.empty((nf,) + tuple(c.dst_shape), device="cuda")
        rc = B.ezsint_batch(z, zin, nf); torch.cuda.synchronize()
        print("dmax", dmax, c.name, "nf", nf, "rc", rc, "path", B.last_path(), "finite", bool(torch.isfinite(z).all()))
    for extrap in (None, "value", "cubic"):
        c = configs.c3(B, 16, "cubic", extrap=extrap); c.defset()
        u, v = c.source_uv(0)
        du, dv = torch.from_numpy(u).cuda(), torch.from_numpy(v).cuda()
        uo, vo = torch.empty(c.dst_shape, device="cuda"), torch.empty(c.dst_shape, device="cuda")
        rc = B.ezuvint(uo, vo, du, dv); torch.cuda.synchronize()
        print("dmax", dmax, c.name, extrap, "rc", rc, "path", B.last_path(), "finite", bool(torch.isfinite(uo).all() and torch.isfinite(vo).all()))
        nj, ni = c.dst_shape
        B.set_rows(7, 45)
        ub, vb = torch.empty((38, ni), device="cuda"), torch.empty((38, ni), device="cuda")
        rc = B.ezuvint(ub, vb, du, dv); torch.cuda.synchronize()
        B.set_rows()
        print("   band rows 7..45 rc", rc, "equal to the full call:", bool(torch.equal(ub, uo[7:45]) and torch.equal(vb, vo[7:45])))
B.set_direct_max_fields(4)
for arith in (0, 1):
    B.set_arith(arith)
    c = configs.c2(B, 3.0, 2.5, "cubic"); c.defset()
    z = np.zeros(c.dst_shape, np.float32)
    rc = B.ezsint(z, c.source(0))
    print("arith", arith, c.name, "rc", rc, "path", B.last_path(), "range", float(z.min()), float(z.max()))
B.set_arith(2)
