# Round-end measurement sweep (run under gpurun): bench lines for every workload + reference arm, launch list, ncu per hot kernel, sanitizer.
# usage: bash scripts/final_sweep.sh <tag>
T=${1:-r2x}
timeout 600 python -X faulthandler -m pytest tests -m gpu -q > gpurun_out/${T}_tests.txt 2>&1; tail -2 gpurun_out/${T}_tests.txt
timeout 400 python bench.py 2>gpurun_out/${T}_bench_c2.err | tail -1 > gpurun_out/${T}_bench_c2.json
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 2>/dev/null | tail -1 > gpurun_out/${T}_bench_ref.json
for w in c1 c3 c4 c5; do timeout 200 python bench.py --no-cpu-baseline --workload $w --steps 20 2>/dev/null | tail -1 > gpurun_out/${T}_bench_$w.json; done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${T}_launches_c2_bench.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extras > /dev/null 2>&1
timeout 400 ncu --set full --clock-control none -k regex:"k_interp_(tiled|rowsep)" -s 4 -c 2 -f -o gpurun_out/${T}_c2_tilekernels python scripts/prof_run.py c2 2 80 4 2>&1 | tail -1
timeout 400 ncu --set full --clock-control none -k regex:rowsep -s 2 -c 1 -f -o gpurun_out/${T}_c5_rowsep python scripts/prof_run.py c5 2 80 4 2>&1 | tail -1
timeout 300 ncu --set full --clock-control none -k regex:k_interp_direct -s 2 -c 1 -f -o gpurun_out/${T}_c3_direct python scripts/prof_run.py c3 2 1 4 2>&1 | tail -1
timeout 300 ncu --set full --clock-control none -k regex:k_interp_direct -s 2 -c 1 -f -o gpurun_out/${T}_c1_direct python scripts/prof_run.py c1 2 1 4 2>&1 | tail -1
python scripts/post_sweep.py $T; rm -f gpurun_out/*.ncu-rep
for t in memcheck racecheck initcheck; do timeout 300 compute-sanitizer --tool $t python scripts/sanitize_small.py > gpurun_out/${T}_sanitizer_$t.txt 2>&1; tail -1 gpurun_out/${T}_sanitizer_$t.txt; done
python - <<PY
import json
for w in ("c2","c5","c1","c3","c4"):
    try:
        d=json.load(open("gpurun_out/${T}_bench_%s.json"%w))
        r=d["roofline"]
        print(w, "%.4f ms"%d["ms_per_step"], "%.3e pts/s"%d["value"], "e2e %.3e"%d["e2e"]["value"], r["kernel"], "frac %.3f"%r["frac"], "step frac %.3f"%r.get("step",{}).get("frac",r["frac"]), "build %.0f ms"%d["gridset_build_ms"], "parity", d.get("parity",{}).get("ok"))
    except Exception as e: print(w, "ERR", e)
d=json.load(open("gpurun_out/${T}_bench_ref.json")); print("ref", d["value"], d["cpu_baseline"]["sample"])
PY
