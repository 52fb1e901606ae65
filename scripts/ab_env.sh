# A/B of run-time switches: bash scripts/ab_env.sh "EZB200_WIN_PERSIST=0" "EZB200_WIN_PERSIST=2" ...
for e in "$@"; do for w in ${WL:-c2}; do echo "env '$e' $w: $(env $e python bench.py --no-cpu-baseline --no-extras --workload $w --steps 20 2>/dev/null | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read()); r=d['roofline']
k=r.get('kernels') or {}
print('%.3f ms  %.3e pts/s step-frac %.3f | '%(d['ms_per_step'],d['value'],r.get('step',r)['frac']) + ' '.join('%s=%.3f'%(n.split('<')[0][9:],x['ms']) for n,x in k.items()))")"; done; done
