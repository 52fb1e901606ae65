# round 2, first GPU check of the re-designed window kernel: parity suite, A/B against the round-1 library, ncu capture
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r2a_pytest.txt
cat gpurun_out/r2a_pytest.txt
for v in "" _r1; do for w in c2 c5; do
  EZB200_LIB=$PWD/librmn_b200/libezb200$v.so timeout 600 python bench.py --no-cpu-baseline --workload $w --steps 20 2>gpurun_out/r2a_err$v$w.txt | tail -1 > gpurun_out/r2a_bench${v}_$w.json
  python -c "
import json,sys
d=json.load(open('gpurun_out/r2a_bench${v}_$w.json'))
print('variant [$v] $w: %.3f ms/step %.3e pts/s step-frac %.3f'%(d['ms_per_step'],d['value'],d['roofline'].get('step',d['roofline'])['frac']), d['roofline'].get('kernels'))
"
done; done
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_interp_(tiled|rowsep)" -s 4 -c 2 -f -o gpurun_out/r2a_c2_tilekernels python scripts/prof_run.py c2 2 80 4 2>&1 | tail -2
