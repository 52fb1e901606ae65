"""First GPU parity probe: product (libezb200.so) vs oracle (libezref.so) on reduced configs. Prints stats."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import librmn_b200
from librmn_b200 import EzLib, configs

B = librmn_b200.load()
R = EzLib(os.path.join(os.path.dirname(librmn_b200.__file__), "..", "oracle", "_ref", "libezref.so"))
print(B.version())

def run(lib, c, nf=1):
    c.defset()
    if c.vector:
        u, v = c.source_uv()
        uo = np.zeros(c.dst_shape, np.float32); vo = np.zeros_like(uo)
        rc = lib.ezuvint(uo, vo, u, v)
        return rc, np.stack([uo, vo])
    z = c.source()
    out = np.zeros(c.dst_shape, np.float32)
    rc = lib.ezsint(out, z)
    return rc, out

def stats(name, ob, orf, rcb, rcr, rng):
    d = np.abs(ob.astype(np.float64) - orf.astype(np.float64))
    nbit = int((ob.view(np.uint32) != orf.view(np.uint32)).sum())
    print("%-44s rc %d/%d  max|d|/range %.3e  bit-mismatch %d / %d  (>1e-5: %d)" % (name, rcb, rcr, d.max() / rng, nbit, ob.size, int((d > 1e-5 * rng).sum())), flush=True)

cases = []
for arith in (0, 1):
    B.set_arith(arith)
    for deg in ("nearest", "linear", "cubic"):
        for mk in (lambda L: configs.c1(L, n=400, degree=deg),
                   lambda L: configs.c2(L, res_src=1.5, res_dst=1.0, degree=deg),
                   lambda L: configs.c3(L, scale=8, degree=deg),
                   lambda L: configs.c3(L, scale=8, degree=deg, extrap="linear"),
                   lambda L: configs.c4(L, ni=160, nj=80, npts=20000, degree=deg),
                   lambda L: configs.c5(L, res_src=1.5, res_dst=1.0, degree=deg)):
            cb, cr = mk(B), mk(R)
            t = time.time(); rcb, ob = run(B, cb); tb = time.time() - t
            rcr, orf = run(R, cr)
            src = cb.source()
            stats("a%d %s %s" % (arith, cb.name, cb.opts.get("extrap_degree", "")), ob, orf, rcb, rcr, float(src.max() - src.min()))
    # scalar on the C3 grids too (DEHORS with max / min / value)
    for ex in ("maximum", "minimum", "value", "nearest", "cubic"):
        cb, cr = configs.c3(B, scale=8, extrap=ex), configs.c3(R, scale=8, extrap=ex)
        cb.vector = cr.vector = False
        cb.opts["extrap_value"] = cr.opts["extrap_value"] = -5.0
        rcb, ob = run(B, cb); rcr, orf = run(R, cr)
        src = cb.source()
        stats("a%d C3 scalar extrap=%s" % (arith, ex), ob, orf, rcb, rcr, float(src.max() - src.min()))
print("launches", B.launch_count())
