"""Two steps around scripts/final_sweep.sh <tag>:
  on the GPU box (called by the sweep itself):  python scripts/post_sweep.py <tag>
      every gpurun_out/<tag>_*.ncu-rep -> gpurun_out/<tag>_ncu_*.txt (per-kernel summary) + gpurun_out/<tag>_traffic_*.json (DRAM bytes
      per launch); the sweep then deletes the reports (gpurun brings back at most 64 MiB)
  here, afterwards:                             python scripts/post_sweep.py <tag> --publish
      gpurun_out/<tag>_*.{json,csv,txt} -> profiles/ ; <tag>_traffic_cN.json -> profiles/traffic_cN.json"""
import csv, json, os, re, shutil, subprocess, sys
T = sys.argv[1]
G, P = "gpurun_out", "profiles"
if "--publish" in sys.argv:
    for f in sorted(os.listdir(G)):
        if f.startswith(T + "_") and (f.endswith(".json") or f.endswith(".csv") or f.endswith(".txt")) and os.path.getsize(os.path.join(G, f)) > 0:
            m = re.match(re.escape(T) + r"_(traffic_c\d\.json)$", f)
            shutil.copy(os.path.join(G, f), os.path.join(P, m.group(1) if m else f))
    sys.exit(0)
def raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    h = rows[0]
    return [dict(zip(h, r)) for r in rows[2:]]
def short(name):
    m = re.search(r"(k_interp_\w+)<(\d+)", name)
    return "%s<%s>" % (m.group(1), m.group(2)) if m else name
for rep, wl, lev, alg in (("c2_tilekernels", "c2", 80, 2888356240), ("c5_rowsep", "c5", 80, 3055088400), ("c3_direct", "c3", 1, None), ("c1_direct", "c1", 1, None)):
    path = os.path.join(G, "%s_%s.ncu-rep" % (T, rep))
    if not os.path.exists(path):
        print("missing", path); continue
    txt = subprocess.run([sys.executable, "scripts/ncu_summary.py", path], capture_output=True, text=True).stdout
    open(os.path.join(G, "%s_ncu_%s.txt" % (T, rep)), "w").write(txt)
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    h, u = rows[0], rows[1]
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    kern = {}
    for r in rows[2:]:
        n = short(r[h.index("Kernel Name")])
        rd = float(r[h.index("dram__bytes_read.sum")]) * scale[u[h.index("dram__bytes_read.sum")]]
        wr = float(r[h.index("dram__bytes_write.sum")]) * scale[u[h.index("dram__bytes_write.sum")]]
        kern[n] = {"dram_bytes_read_per_launch": rd, "dram_bytes_write_per_launch": wr, "dram_bytes_per_launch": rd + wr,
                   "gpu_time_us": float(r[h.index("gpu__time_duration.sum")]) * {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(u[h.index("gpu__time_duration.sum")], 1.0)}
    tj = {"workload": wl, "levels": lev,
          "source": "ncu --set full --clock-control none, one launch per kernel of one %d-level step: profiles/%s_ncu_%s.txt (dram__bytes_read.sum + dram__bytes_write.sum)" % (lev, T, rep),
          "kernels": kern, "dram_bytes_per_launch": sum(k["dram_bytes_per_launch"] for k in kern.values())}
    if alg:
        tj["algorithmic_bytes_per_step"] = alg
    json.dump(tj, open(os.path.join(G, "%s_traffic_%s.json" % (T, wl)), "w"), indent=1)
    print(wl, {k: "%.0f MB in %.1f us" % (v["dram_bytes_per_launch"] / 1e6, v["gpu_time_us"]) for k, v in kern.items()})
