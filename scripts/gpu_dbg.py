import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import librmn_b200
from librmn_b200 import EzLib, configs
B = librmn_b200.load()
R = EzLib(os.path.join(os.path.dirname(librmn_b200.__file__), "..", "oracle", "_ref", "libezref.so"))
B.set_arith(0)
cb, cr = configs.c5(B, 1.5, 1.0, "nearest"), configs.c5(R, 1.5, 1.0, "nearest")
outs = []
for lib, c in ((B, cb), (R, cr)):
    c.defset(); z = c.source(); o = np.zeros(c.dst_shape, np.float32); print(lib.ezsint(o, z)); outs.append(o)
d = outs[0] != outs[1]
jj, ii = np.nonzero(d)
print("rows", np.unique(jj), "cols", len(np.unique(ii)))
print(outs[0][jj[0], :5], outs[1][jj[0], :5])
x, y, zb, cnt = B.get_gridset(cb.gdout, cb.gdin, cb.n_out)
print("counts", cnt)
y = y.reshape(cb.dst_shape); zb = zb.reshape(cb.dst_shape)
for j in np.unique(jj): print(j, y[j, :3], zb[j, :3])
print('row0', y[0,:3], zb[0,:3], 'rowlast', y[-1,:3], zb[-1,:3])
