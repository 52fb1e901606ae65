import sys, os, numpy as np
ROOT=os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0,ROOT); sys.path.insert(0,os.path.join(ROOT,'tests'))
import librmn_b200
from librmn_b200 import EzLib
import cases_extra as ce, cases
B=librmn_b200.load(); R=EzLib(os.path.join(ROOT,'oracle','_ref','libezref.so'))
deg=sys.argv[1] if len(sys.argv)>1 else 'nearest'
want=sys.argv[2:] or ['A_north_to_L_%s_winds'%deg]
for cb,cr in zip(ce.regular_source_cases(B,deg), ce.regular_source_cases(R,deg)):
    if cb.name not in want: continue
    rcb,ob=ce.run(B,cb); rcr,orf=ce.run(R,cr,expanded=hasattr(cr,'equiv'))
    rng=cases.field_range(cb)
    d=np.abs(ob.astype(np.float64)-orf)/rng
    print(cb.name, 'rc',rcb,rcr,'max',d.max())
    for comp in range(2):
        rows=np.where(d[comp].max(axis=1)>1e-5)[0]
        print(' comp',comp,'bad rows',rows, 'of', d.shape[1], 'max per bad row', np.round(d[comp].max(axis=1)[rows],4))
    x,y,zb,counts=B.get_gridset(cb.gdout,cb.gdin,cb.n_out)
    print(' counts',counts, 'y rows', y.reshape(cb.dst_shape)[:4,0], y.reshape(cb.dst_shape)[-4:,0])
