# single-call workloads (C1, C3): direct kernel variants (EZB200_LIB) vs tile kernels (EZB200_DIRECT_MAX_FIELDS=0); device-resident ms per call
run() { env $1 timeout 200 python bench.py --workload $2 --no-cpu-baseline --no-extras --steps 20 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('$2 [$1]: %.4f ms/step  step-frac %.3f  kernels %s'%(d['ms_per_step'], r.get('step',{}).get('frac',r['frac']), {k:round(v.get('ms',0),4) for k,v in r.get('kernels',{}).items()}))"; }
for w in c1 c3; do
  run "EZB200_DIRECT_MAX_FIELDS=0" $w
  run "EZB200_DIRECT_MAX_FIELDS=4" $w
  for v in va vb vc vd; do run "EZB200_LIB=librmn_b200/libezb200_$v.so" $w; done
done
