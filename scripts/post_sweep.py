"""After scripts/final_sweep.sh <tag> (run under gpurun): turn gpurun_out/<tag>_* into the tracked summaries under profiles/
(ncu per-kernel text summaries, bench lines, launch list, sanitizer logs) and refresh profiles/traffic_c{1,2,3,5}.json from the
ncu captures.    usage: python scripts/post_sweep.py <tag>"""
import csv, json, os, re, shutil, subprocess, sys
T = sys.argv[1]
G, P = "gpurun_out", "profiles"
for f in sorted(os.listdir(G)):
    if f.startswith(T + "_") and (f.endswith(".json") or f.endswith(".csv") or f.endswith(".txt")) and os.path.getsize(os.path.join(G, f)) > 0:
        shutil.copy(os.path.join(G, f), os.path.join(P, f))
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
    open(os.path.join(P, "%s_ncu_%s.txt" % (T, rep)), "w").write(txt)
    ks = {}
    for r in raw(path):
        rd, wr = float(r["dram__bytes_read.sum"]), float(r["dram__bytes_write.sum"])
        # units: ncu prints Mbyte / Gbyte per column; the raw page keeps one unit per column, given in row 1 -- re-read it
        ks[short(r["Kernel Name"])] = (rd, wr)
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
    json.dump(tj, open(os.path.join(P, "traffic_%s.json" % wl), "w"), indent=1)
    print(wl, {k: "%.0f MB in %.1f us" % (v["dram_bytes_per_launch"] / 1e6, v["gpu_time_us"]) for k, v in kern.items()})
