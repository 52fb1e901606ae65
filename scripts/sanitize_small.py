"""Small tiled-path run for compute-sanitizer (memcheck / racecheck): row-separable + window + list kernels, tail stages."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import librmn_b200
from librmn_b200 import configs
B = librmn_b200.load()
B.set_arith(2)
for mk, nf in ((lambda: configs.c2(B, 3.0, 2.0, "cubic"), 6), (lambda: configs.c5(B, 3.0, 2.0, "cubic"), 5), (lambda: configs.c1(B, 96, "linear"), 3)):
    c = mk(); c.defset()
    zin = torch.from_numpy(np.stack([c.source(k) for k in range(nf)])).cuda().contiguous()
    z = torch.empty((nf,) + tuple(c.dst_shape), device="cuda")
    rc = B.ezsint_batch(z, zin, nf); torch.cuda.synchronize()
    print(c.name, "rc", rc, "path", B.last_path(), "finite", bool(torch.isfinite(z).all()))
