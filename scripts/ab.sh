# A/B of experiment builds (librmn_b200/build.py --out=...): WL="c2 c5" bash scripts/ab.sh "" _m3 _p1 ...
# prints per variant and workload: step ms, points/s, step roofline fraction and the per-kernel milliseconds
for v in "$@"; do for w in ${WL:-c2}; do echo "variant '$v' $w: $(EZB200_LIB=$PWD/librmn_b200/libezb200$v.so python bench.py --no-cpu-baseline --no-extras --workload $w --steps 20 2>/dev/null | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read()); r=d['roofline']
k=r.get('kernels') or {}
print('%.3f ms  %.3e pts/s step-frac %.3f | '%(d['ms_per_step'],d['value'],r.get('step',r)['frac']) + ' '.join('%s=%.3f'%(n.split('<')[0][9:],x['ms']) for n,x in k.items()))")"; done; done
