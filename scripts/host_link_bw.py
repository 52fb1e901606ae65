"""Host <-> device link bound of a multi-GPU box, independent of the library (VERDICT r1 next #4d).

For N = 1, 2, 4, 8 (as many GPUs as are visible): every device simultaneously reads `in_mb` MB from and writes `out_mb` MB to pinned
host memory (the C2 step's ratio: 755 MB in, 2075 MB out, split over the N devices), one thread per device, copies on two
streams per device (full duplex).  Two placements of the pinned buffers:
  plain      : allocated by the main thread wherever the OS puts them (what bench.py's e2e leg and a typical caller do)
  numa_local : each device's block allocated and first-touched by a thread pinned to the CPUs of that GPU's NUMA node
               (/sys/bus/pci/devices/<bdf>/numa_node, /sys/devices/system/node/node<k>/cpulist)
Prints one JSON line: aggregate GB/s per N and placement.
"""
import json, os, threading, time
import torch


def gpu_numa_node(dev):
    bdf = None
    if True:
        try:
            import subprocess
            out = subprocess.run(["nvidia-smi", "-i", str(dev), "--query-gpu=pci.bus_id", "--format=csv,noheader"], capture_output=True, text=True).stdout.strip()
            bdf = out
        except Exception:
            return None
    bdf = bdf.lower()
    if len(bdf.split(":")[0]) == 8:
        bdf = bdf[4:]
    try:
        with open("/sys/bus/pci/devices/%s/numa_node" % bdf) as f:
            return int(f.read().strip())
    except Exception:
        return None


def node_cpus(node):
    try:
        with open("/sys/devices/system/node/node%d/cpulist" % node) as f:
            txt = f.read().strip()
        cpus = []
        for part in txt.split(","):
            a, _, b = part.partition("-")
            cpus += list(range(int(a), int(b or a) + 1))
        return cpus
    except Exception:
        return None


def run(ndev, local, total_in_mb=755, total_out_mb=2075, reps=5):
    in_n = int(total_in_mb * 1e6 / 4 / ndev)
    out_n = int(total_out_mb * 1e6 / 4 / ndev)
    bufs = [None] * ndev

    def alloc(d):
        if local:
            node = gpu_numa_node(d)
            cpus = node_cpus(node) if node is not None and node >= 0 else None
            if cpus:
                try:
                    os.sched_setaffinity(0, cpus)
                except Exception:
                    pass
        torch.cuda.set_device(d)
        hi = torch.empty(in_n, dtype=torch.float32).pin_memory()
        ho = torch.empty(out_n, dtype=torch.float32).pin_memory()
        hi.fill_(1.0); ho.fill_(0.0)   # first touch
        di = torch.empty(in_n, dtype=torch.float32, device="cuda:%d" % d)
        do = torch.ones(out_n, dtype=torch.float32, device="cuda:%d" % d)
        bufs[d] = (hi, ho, di, do, torch.cuda.Stream(device=d), torch.cuda.Stream(device=d))

    th = [threading.Thread(target=alloc, args=(d,)) for d in range(ndev)]
    [t.start() for t in th]; [t.join() for t in th]
    barrier = threading.Barrier(ndev + 1)

    def work(d):
        torch.cuda.set_device(d)
        hi, ho, di, do, s1, s2 = bufs[d]
        for rep in range(reps + 1):
            barrier.wait()
            with torch.cuda.stream(s1):
                di.copy_(hi, non_blocking=True)
            with torch.cuda.stream(s2):
                ho.copy_(do, non_blocking=True)
            s1.synchronize(); s2.synchronize()
            barrier.wait()

    th = [threading.Thread(target=work, args=(d,)) for d in range(ndev)]
    [t.start() for t in th]
    times = []
    for rep in range(reps + 1):
        barrier.wait()
        t0 = time.perf_counter()
        barrier.wait()
        times.append(time.perf_counter() - t0)
    [t.join() for t in th]
    best = min(times[1:])
    return (total_in_mb + total_out_mb) / 1e3 / best, best * 1e3


if __name__ == "__main__":
    nvis = torch.cuda.device_count()
    res = {"gpus_visible": nvis, "numa_nodes_of_gpus": [gpu_numa_node(d) for d in range(nvis)], "host_cpus": os.cpu_count(), "runs": {}}
    for n in (1, 2, 4, 8):
        if n > nvis:
            break
        for local in (False, True):
            gbs, ms = run(n, local)
            res["runs"]["%d_gpus_%s" % (n, "numa_local" if local else "plain")] = {"aggregate_GBps": round(gbs, 1), "ms_for_one_C2_step_of_bytes": round(ms, 2)}
    print(json.dumps(res))
