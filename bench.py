#!/usr/bin/env python
"""bench.py -- headline benchmark of the EZ interpolator hot path (BASELINE.json: output points/s of
c_ezsint / c_ezuvint, and achieved HBM GB/s against the measured peak).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload c2|c5|c1|c3|c4]

A "step" is ONE pass of the hot path over one batch of synthetic fields: c_ezsint_batch(zout, zin, L) for
the configured workload (default = BASELINE.json configs[1]: GEM Yin-Yang 'U' -> 0.1 deg 'L', cubic, 80 levels).
  value : output points/s with source and target fields resident in HBM (device pointers through the C ABI)
  e2e   : the same call with HOST (pinned) buffers -- host->device and device->host copies inside the timed region
          (e2e.pageable: the same with plain malloc'ed arrays, what an unmodified caller hands over)
N > 1 (torchrun, one rank per GPU): levels are partitioned, every rank holds a replicated gridset and its own
L levels (weak scaling, no data-path collective); time = max over ranks.

Beside the contract's K timed steps the line carries
  sustained : the same step looped for >= 1 s (clock samples are taken over this region too)
  parity    : level 0 of this very run compared with the oracle, outside the timed region
  extra     : sub-records for c_ezuvint (C3), C5 and the fp64 weight-form arm of C2 (--arith 1), each with value / roofline / e2e
  collective: (N > 1, --gather) an NCCL all-gather of the ranks' level blocks, timed beside the kernel time

`--impl reference` times the reference's own CPU implementation (oracle/_ref/libezref.so: its unmodified
src/interp/*.c + our restatement of its Fortran leaves) on the host cores, on a bounded sample of the same
workload.  The oracle is never on the product path.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

REF_SO = os.path.join(ROOT, "oracle", "_ref", "libezref.so")
REF_VEC_SO = os.path.join(ROOT, "oracle", "_ref", "libezref_vec.so")   # same sources, the reference's default flags (oracle/Makefile)


# --------------------------------------------------------------------------------------------------
def workload(lib, name):
    from librmn_b200 import configs
    if name == "c2":
        return configs.c2(lib, 0.15, 0.1, "cubic"), 80, "c_ezsint cubic, GEM Yin-Yang 'U' (2 x 1841x641 Z/E panels) -> 'L' 0.1deg 3600x1801, 80 levels"
    if name == "c5":
        return configs.c5(lib, 0.15, 0.1, "cubic"), 80, "c_ezsint cubic, global 'Z' 0.15deg 2401x1200 -> 'L' 0.1deg 3600x1801, 80 levels per GPU"
    if name == "c1":
        return configs.c1(lib, 2048, "linear"), 1, "c_ezsint bilinear, 'L' 720x361 -> 'N' 2048x2048, 1 field"
    if name == "c3":
        return configs.c3(lib, 1, "cubic"), 1, "c_ezuvint cubic, rotated 'Z' 2576x1456 -> 'L' 3200x1600, 1 (u,v) pair"
    if name == "c4":
        return configs.c4(lib, 1600, 800, 1_000_000, "linear"), 16, "c_ezsint linear, 'G' 1600x800 -> 'Y' 1M points, 16 fields"
    raise SystemExit("unknown workload %s" % name)


def algorithmic_bytes(case, nfields):
    """F*(4 N_src + 4 N_out) + 9 N_out per batch (x, y, zone byte read once per batch)."""
    comps = 2 if case.vector else 1
    return comps * nfields * 4 * (case.n_src + case.n_out) + 9 * case.n_out


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    return 6650.0, "fallback"


class ClockSampler(threading.Thread):
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.rows = []
        self.stop_flag = False

    def run(self):
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([c.strip() for c in out.split(",")])
            except Exception:
                pass
            time.sleep(0.1)

    def summary(self):
        sm = [float(r[0]) for r in self.rows if r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in self.rows for n, v in zip(names, r[3:7]) if v.lower().startswith("active")})
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(self.rows)}


# --------------------------------------------------------------------------------------------------
def cpu_reference_run(name, nfields_sample, nthreads, steps, warmup, so=REF_SO):
    """Times oracle/_ref (the reference's C + restated Fortran leaves) on host cores. Returns (pts/s, ms/step, case, desc)."""
    import ctypes as C
    from librmn_b200 import EzLib
    if not os.path.exists(so):
        raise SystemExit("%s missing: run `make -C oracle` where /root/reference exists" % so)
    R = EzLib(so)
    case, _, desc = workload(R, name)
    case.defset()
    nin, nout = case.n_src, case.n_out
    zin = np.stack([case.source(k).ravel() for k in range(nfields_sample)])
    zout = np.zeros((nfields_sample, nout), np.float32)
    vin = vout = None
    if case.vector:
        uv = [case.source_uv(k) for k in range(nfields_sample)]
        zin = np.stack([u.ravel() for u, _ in uv])
        vin = np.stack([v.ravel() for _, v in uv])
        vout = np.zeros_like(zout)
    # one warm single-threaded call builds the gridset (the reference's registry is not thread safe while building)
    if case.vector:
        R.ezuvint(zout[0], vout[0], zin[0], vin[0])
    else:
        R.ezsint(zout[0], zin[0])
    fn = R.lib.ezref_mt
    fn.argtypes = [C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_int32, C.c_int32, C.c_char_p, C.c_char_p]
    interp = case.opts.get("interp_degree", "cubic").encode()
    extrap = case.opts.get("extrap_degree", "maximum").encode()

    def step():
        return fn(case.gdout, case.gdin, zout.ctypes.data, zin.ctypes.data, vout.ctypes.data if vout is not None else None,
                  vin.ctypes.data if vin is not None else None, nin, nout, nfields_sample, nthreads, interp, extrap)

    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = (time.perf_counter() - t0) / steps
    return nout * nfields_sample / dt, dt * 1e3, case, desc


def host_can_run_vec_build():
    try:
        with open("/proc/cpuinfo") as f:
            for ln in f:
                if ln.startswith("flags"):
                    fl = set(ln.split(":", 1)[1].split())
                    return {"avx2", "fma", "bmi2", "movbe"} <= fl
    except Exception:
        pass
    return False


def oracle_level(name):
    """Level 0 of the workload through the oracle (host arrays). Returns (rc, stacked outputs, range of the source)."""
    from librmn_b200 import EzLib
    R = EzLib(REF_SO)
    case, _, _ = workload(R, name)
    case.defset()
    if case.vector:
        u, v = case.source_uv(0)
        uo, vo = np.zeros(case.dst_shape, np.float32), np.zeros(case.dst_shape, np.float32)
        rc = R.ezuvint(uo, vo, u, v)
        return rc, np.stack([uo.ravel(), vo.ravel()]), float(max(u.max() - u.min(), v.max() - v.min()))
    z = case.source(0)
    o = np.zeros(case.dst_shape, np.float32)
    rc = R.ezsint(o, z)
    return rc, o.ravel()[None], float(z.max() - z.min())


# --------------------------------------------------------------------------------------------------
def measure(B, name, arith, steps, warmup, levels, rank, world, stream, sampler_cb=None, sustain_s=1.0, pageable=True, parity=True,
            gather=False):
    """One workload through the C ABI: device-resident K steps, sustained loop, per-kernel split, e2e (pinned, pageable),
    parity of level 0 against the oracle.  Returns a dict of sub-records."""
    import torch
    import torch.distributed as dist

    t_phase = [time.perf_counter()]

    def phase(what):   # progress on stderr (the JSON line is the only thing on stdout)
        now = time.perf_counter()
        if rank == 0:
            sys.stderr.write("[bench %s arith=%d] %-28s %7.2f s\n" % (name, arith, what, now - t_phase[0]))
            sys.stderr.flush()
        t_phase[0] = now

    B.set_arith(arith)
    case, nlev, desc = workload(B, name)
    if levels:
        nlev = levels
    nin, nout = case.n_src, case.n_out
    case.defset()
    # fewer fields than GPUs (single-call workloads C1 / C3 / C4 at N > 1): the target ROWS are split over the ranks instead of
    # the levels; every rank holds the whole source field and produces its band of whole tile rows
    band_mode = world > 1 and nlev < world and case.dst_shape[0] >= 32 * world
    nout_full = nout
    if band_mode:
        from librmn_b200 import sharding
        j0, j1 = sharding.row_band(case.dst_shape[0], rank, world)
        B.set_rows(j0, j1)
        nout = (j1 - j0) * case.dst_shape[1]
    work_scale = (1.0 / world) if band_mode else 1.0   # points per step summed over the ranks = world * nout_full * nlev * work_scale

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # synthetic analytic fields, generated once on the host and made resident in HBM before the timed region.
    # every rank owns its own nlev levels (level k of rank r = base field + 0.01 (r nlev + k); interpolation is linear in the field).
    comps = 2 if case.vector else 1
    h_in = [torch.empty((nlev, nin), dtype=torch.float32, pin_memory=True) for _ in range(comps)]
    base = [case.source_uv(0) if case.vector else (case.source(0),)][0]
    for c in range(comps):
        b = torch.from_numpy(np.ascontiguousarray(base[c]).ravel())
        for k in range(nlev):
            h_in[c][k].copy_(b)
            if rank * nlev + k:
                h_in[c][k].add_(0.01 * (rank * nlev + k))
    d_in = [h.cuda(non_blocking=True) for h in h_in]
    d_out = [torch.empty((nlev, nout), dtype=torch.float32, device="cuda") for _ in range(comps)]
    h_out = [torch.empty((nlev, nout), dtype=torch.float32, pin_memory=True) for _ in range(comps)]
    torch.cuda.synchronize()
    phase("fields generated + resident")

    def step_dev():
        if case.vector:
            return B.ezuvint_batch(d_out[0], d_out[1], d_in[0], d_in[1], nlev)
        return B.ezsint_batch(d_out[0], d_in[0], nlev)

    def step_host(hi, ho):
        if case.vector:
            return B.ezuvint_batch(ho[0], ho[1], hi[0], hi[1], nlev)
        return B.ezsint_batch(ho[0], hi[0], nlev)

    torch.cuda.synchronize()
    t_build = time.perf_counter()
    rc = step_dev()   # builds the gridset and the tile plan (one-time K12-K15 cost, reported separately as gridset_build_ms)
    if rc < 0:
        raise SystemExit("hot path failed rc=%d (%s)" % (rc, name))
    torch.cuda.synchronize()
    gridset_build_ms = (time.perf_counter() - t_build) * 1e3   # first call = build + one step
    for _ in range(max(3, warmup)):
        step_dev()
    torch.cuda.synchronize()
    phase("gridset build + warm-up")
    if sampler_cb:
        sampler_cb()   # starts the clock sampler: it runs over the timed, the sustained and the e2e regions

    # ---- device-resident timing: EXACTLY K steps, CUDA events on the launching stream ----
    # L2 rule: a step whose fields fit the 126 MB L2 is timed with an L2 flush (a 256 MB write) between iterations, outside
    # the per-step event pairs, and its time is the sum of the per-step intervals; larger steps evict themselves.
    footprint = comps * nlev * (nin + nout) * 4
    flush_buf = torch.empty(64 * 1024 * 1024, dtype=torch.float32, device="cuda") if footprint < 300e6 else None
    l2_note = ("L2 flushed between timed iterations (256 MB write; step footprint %.0f MB)" % (footprint / 1e6)) if flush_buf is not None else (
        "inputs larger than L2 (%.0f MB in + %.0f MB out per step)" % (comps * nlev * nin * 4 / 1e6, comps * nlev * nout * 4 / 1e6))
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    B.launch_count(reset=True)
    barrier()
    e_all0, e_all1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e_all0.record(stream)
    for a, b in evs:
        if flush_buf is not None:
            flush_buf.zero_()
        a.record(stream)
        step_dev()
        b.record(stream)
    e_all1.record(stream)
    barrier()
    launches = B.launch_count()
    path, nlist = B.last_path()
    step_ms = [a.elapsed_time(b) for a, b in evs]
    total_ms = sum(step_ms) if flush_buf is not None else e_all0.elapsed_time(e_all1)
    t = torch.tensor([total_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_per_step = float(t.item()) / steps
    value = world * nout_full * work_scale * nlev / (ms_per_step * 1e-3)

    phase("timed steps")
    # ---- sustained: the same step back to back for >= sustain_s seconds (the K-step region above is a burst of tens of ms) ----
    sustained = None
    if sustain_s > 0:
        n_s = max(steps, int(sustain_s * 1e3 / max(ms_per_step, 1e-3)) + 1)
        n_s = min(n_s, 20000)
        barrier()
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s0.record(stream)
        for _ in range(n_s):   # (back to back, no flush: for steps that fit the L2 this is the L2-warm rate and is labelled so)
            step_dev()
        s1.record(stream)
        barrier()
        ts = torch.tensor([s0.elapsed_time(s1)], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(ts, op=dist.ReduceOp.MAX)
        s_ms = float(ts.item()) / n_s
        sustained = {"steps": n_s, "seconds": float(ts.item()) * 1e-3, "ms_per_step": s_ms, "value": world * nout_full * work_scale * nlev / (s_ms * 1e-3), "unit": "points/s",
                     "l2": "warm" if flush_buf is not None else "evicted by the step itself"}

    phase("sustained loop")
    # per-kernel durations of the same step, measured live with CUDA events recorded by the library on the launching
    # stream around each kernel (c_ezb200_set_timing): a few extra steps outside the K timed ones
    B.set_timing(True)
    kms = []
    tiles3 = (0, 0, 0)
    for _ in range(5):
        step_dev()
        ms3, tiles3 = B.last_kernel_ms()
        if ms3:
            kms.append(ms3)
    B.set_timing(False)
    torch.cuda.synchronize()

    # ---- end-to-end: host buffers through the same C-ABI call, copies inside the timed region ----
    def time_host(hi, ho, nsteps):
        for _ in range(2):
            step_host(hi, ho)
        barrier()
        t0 = time.perf_counter()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(nsteps):
            step_host(hi, ho)
        e1.record(stream)
        barrier()
        wall_ms = (time.perf_counter() - t0) * 1e3
        te = torch.tensor([max(e0.elapsed_time(e1), wall_ms)], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        return float(te.item()) / nsteps

    phase("per-kernel timing")
    e2e_steps = max(3, min(steps, 10))
    e2e_ms = time_host(h_in, h_out, e2e_steps)
    phase("e2e pinned")
    e2e = {"value": world * nout_full * work_scale * nlev / (e2e_ms * 1e-3), "unit": "points/s", "h2d_bytes_per_step": comps * nlev * nin * 4,
           "d2h_bytes_per_step": comps * nlev * nout * 4, "ms_per_step": e2e_ms, "host_memory": "pinned"}
    if pageable:
        # what an unmodified caller hands over: plain (malloc) arrays
        p_in = [np.array(h.numpy(), copy=True) for h in h_in]
        p_out = [np.empty((nlev, nout), np.float32) for _ in range(comps)]
        pg_ms = time_host(p_in, p_out, max(2, e2e_steps // 2))
        e2e["pageable"] = {"value": world * nout_full * work_scale * nlev / (pg_ms * 1e-3), "unit": "points/s", "ms_per_step": pg_ms, "host_memory": "pageable (malloc)"}
        same = all(np.array_equal(p_out[c], h_out[c].numpy()) for c in range(comps))
        e2e["pageable"]["equals_pinned_result"] = bool(same)
        del p_in, p_out
        phase("e2e pageable")

    # ---- one process, every visible GPU: the same host-array call fanned out by the library (c_ezb200_set_devices) ----
    ndev_vis = torch.cuda.device_count()
    if world == 1 and ndev_vis > 1 and nlev >= 2 * ndev_vis and not band_mode:
        for nd in sorted({2, ndev_vis}):
            if nd > ndev_vis:
                continue
            used = B.set_devices(nd)
            try:
                fo_ms = time_host(h_in, h_out, e2e_steps)
            finally:
                B.set_devices(1)
            e2e.setdefault("fanout", {})["%d_gpus" % used] = {"value": nout_full * nlev / (fo_ms * 1e-3), "unit": "points/s", "ms_per_step": fo_ms,
                                                              "host_memory": "pinned", "what": "ONE process, one c_ezsint_batch call on host arrays, "
                                                              "levels cut into %d blocks by the library, one gridset replica per device" % used}
        phase("e2e fan-out")

    # ---- optional collective: gather every rank's level block on every rank (NCCL all-gather over NVLink) ----
    collective = None
    if gather and world > 1:
        from librmn_b200 import sharding

        def do_gather():
            if band_mode:
                return sharding.gather_rows(d_out[0].view(nlev, -1, case.dst_shape[1])[0], case.dst_shape[0])
            return sharding.gather_levels(d_out[0], world * nlev)

        for _ in range(2):
            do_gather()
        barrier()
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        g0.record()
        nrep = 5
        for _ in range(nrep):
            full = do_gather()
        g1.record()
        barrier()
        tg = torch.tensor([g0.elapsed_time(g1) / nrep], dtype=torch.float64, device="cuda")
        dist.all_reduce(tg, op=dist.ReduceOp.MAX)
        gms = float(tg.item())
        recv = ((nout_full - nout) if band_mode else (world - 1) * nlev * nout) * 4
        collective = {"op": ("all_gather of %d row bands of one %d-point field (NCCL, padded to the tallest band)" % (world, nout_full)) if band_mode else (
                          "all_gather of %d x %d levels of %d points (NCCL)" % (world, nlev, nout)),
                      "ms": gms, "ms_per_level": gms / (1 if band_mode else world * nlev),
                      "bytes_received_per_rank": recv, "GBps_per_rank": recv / (gms * 1e-3) / 1e9,
                      "kernel_ms_per_step": ms_per_step, "gathered_shape": list(full.shape)}
        del full

    # ---- parity of THIS run's output (level 0 of rank 0) against the oracle, outside every timed region ----
    par = None
    if parity and rank == 0 and os.path.exists(REF_SO):
        try:
            step_dev()
            torch.cuda.synchronize()
            got = np.stack([d_out[c][0].cpu().numpy() for c in range(comps)])
            rc_r, ref, rng = oracle_level(name)
            if band_mode:
                ref = ref.reshape(comps, case.dst_shape[0], case.dst_shape[1])[:, j0:j1].reshape(comps, -1)
            d = np.abs(got.astype(np.float64) - ref) / rng
            # Yin-Yang: the panel choice of a point exactly on the mask border may flip (a handful of points at most)
            nbad = int((d > 1e-5).sum())
            par = {"checked": "level 0 of this run against oracle/_ref/libezref.so, %d values" % got.size, "tol_of_range": 1e-5,
                   "max_err_over_range": float(np.sort(d.ravel())[-9]) if got.size > 16 else float(d.max()), "values_over_tol": nbad,
                   "ok": bool(nbad <= 8), "rc": [int(rc), int(rc_r)]}
        except Exception as ex:
            par = {"checked": "failed: %r" % (ex,), "ok": False}

    phase("collective + parity")
    # ---- roofline of the dominant kernel ----
    peak, peak_kind = measured_peak()
    abytes = comps * nlev * 4 * (nin + nout) + 9 * nout   # SURVEY.md 8(d) for this rank's share (algorithmic_bytes() when nothing is split)
    step_kern_ms = statistics.mean(step_ms)
    deg = {"nearest": 0, "linear": 1, "cubic": 3}[case.opts.get("interp_degree", "cubic")]
    if kms and (path & 10):
        rs_ms, win_ms, list_ms = [statistics.median(k[i] for k in kms) for i in range(3)]
        t_rs, t_win, n_list = tiles3
        tile_pts = max(1, nout - n_list)
        share = {"rowsep": tile_pts * (t_rs / max(1, t_rs + t_win)) / nout, "window": tile_pts * (t_win / max(1, t_rs + t_win)) / nout}
        kernels = {"k_interp_rowsep<%d>" % deg: {"ms": rs_ms, "tiles": t_rs, "points_share": share["rowsep"]},
                   "k_interp_tiled<%d>" % deg: {"ms": win_ms, "tiles": t_win, "points_share": share["window"]},
                   "k_interp_fast (list)": {"ms": list_ms, "points": n_list}}
        for kn, sh in (("k_interp_rowsep<%d>" % deg, share["rowsep"]), ("k_interp_tiled<%d>" % deg, share["window"])):
            if kernels[kn]["ms"] > 0 and sh > 0:
                kernels[kn]["frac"] = abytes * sh / (kernels[kn]["ms"] * 1e-3) / 1e9 / peak
        if rs_ms >= win_ms:
            dom, dom_ms, dom_share = "k_interp_rowsep<%d>" % deg, rs_ms, share["rowsep"]
        else:
            dom, dom_ms, dom_share = "k_interp_tiled<%d>" % deg, win_ms, share["window"]
        dom_bytes = abytes * dom_share
        achieved = dom_bytes / (dom_ms * 1e-3) / 1e9
        roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": None,
                    "peak_kind": peak_kind, "kernel": dom, "kernel_ms": dom_ms, "algorithmic_bytes_per_launch": dom_bytes,
                    "kernels": kernels,
                    "step": {"algorithmic_bytes": abytes, "ms": step_kern_ms, "achieved": abytes / (step_kern_ms * 1e-3) / 1e9,
                             "frac": abytes / (step_kern_ms * 1e-3) / 1e9 / peak},
                    "points_on_per_point_list": nlist}
    elif kms and (path & 16):
        # direct kernel (calls of 1-4 fields): [scalars pre-pass | direct kernel | zone-list kernel], timed in line; in the timed
        # steps the two small kernels run beside the direct kernel on a forked stream
        pre_ms, dir_ms, zl_ms = [statistics.median(k[i] for k in kms) for i in range(3)]
        dom = "k_interp_direct<%d>" % deg
        achieved = abytes / (dir_ms * 1e-3) / 1e9
        roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": None,
                    "peak_kind": peak_kind, "kernel": dom, "kernel_ms": dir_ms, "algorithmic_bytes_per_launch": abytes,
                    "kernels": {"k_field_scalars": {"ms": pre_ms}, dom: {"ms": dir_ms, "frac": achieved / peak}, "k_interp_fast (zone list)": {"ms": zl_ms, "points": nlist}},
                    "step": {"algorithmic_bytes": abytes, "ms": step_kern_ms, "achieved": abytes / (step_kern_ms * 1e-3) / 1e9,
                             "frac": abytes / (step_kern_ms * 1e-3) / 1e9 / peak},
                    "points_on_per_point_list": nlist}
    else:
        achieved = abytes / (step_kern_ms * 1e-3) / 1e9
        roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": None,
                    "peak_kind": peak_kind, "kernel": "k_interp_fast<%d>" % deg, "kernel_ms": step_kern_ms,
                    "algorithmic_bytes_per_launch": abytes, "points_on_per_point_list": nlist}
    roofline["peak_nominal"] = 8000.0                    # fractions against both denominators
    roofline["frac_nominal"] = roofline["achieved"] / 8000.0
    traffic_file = os.path.join(ROOT, "profiles", "traffic_%s.json" % name)
    if os.path.exists(traffic_file):
        with open(traffic_file) as f:
            tj = json.load(f)
        roofline["traffic"] = tj.get("kernels", {}).get(roofline["kernel"], {}).get("dram_bytes_per_launch", tj.get("dram_bytes_per_launch"))
        roofline["traffic_source"] = tj.get("source")

    if band_mode:
        B.set_rows()
    del d_in, d_out, h_in, h_out, flush_buf
    torch.cuda.empty_cache()
    return {"value": value, "ms_per_step": ms_per_step, "desc": desc, "nlev": nlev, "nin": nin, "nout": nout, "comps": comps, "e2e": e2e,
            "launches": int(launches), "gridset_build_ms": gridset_build_ms, "roofline": roofline, "sustained": sustained, "parity": par,
            "collective": collective, "arith": arith, "l2": l2_note, "band_mode": band_mode}


def sub_record(m):
    """Compact form of a measure() result for the `extra` block."""
    r = m["roofline"]
    out = {"workload": m["desc"], "arith": m["arith"], "dtype": "f32" if m["arith"] == 2 else "f64", "value": m["value"], "unit": "points/s",
           "ms_per_step": m["ms_per_step"], "levels": m["nlev"], "gpu_launches": m["launches"], "l2": m["l2"],
           "roofline": {"kernel": r["kernel"], "frac": r["frac"], "achieved": r["achieved"], "peak": r["peak"], "unit": "GB/s",
                        "step_frac": r.get("step", {}).get("frac", r["frac"]),
                        "kernels": {k: {kk: vv for kk, vv in v.items() if kk in ("ms", "frac")} for k, v in (r.get("kernels") or {}).items()}},
           "e2e": m["e2e"], "parity": m["parity"]}
    if m["sustained"]:
        out["sustained"] = m["sustained"]
    return out


# --------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c2")
    ap.add_argument("--levels", type=int, default=0, help="override the number of fields per step")
    ap.add_argument("--cpu-sample-levels", type=int, default=8)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the extra sub-records (other workloads, fp64 arm, pageable e2e, sustained loop, parity)")
    ap.add_argument("--gather", action="store_true", help="N > 1: also time an NCCL all-gather of the ranks' level blocks")
    ap.add_argument("--arith", type=int, default=2)
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    metric = "output_points_per_second"
    unit = "points/s"
    cores = os.cpu_count() or 1

    if args.impl == "reference":
        if rank != 0:
            return 0
        steps = max(1, min(args.steps, 5))
        warm = max(1, min(args.warmup, 2))
        nsample = max(args.cpu_sample_levels, cores)   # at least one level per host thread, so that every core has work
        v, ms, case, desc = cpu_reference_run(args.workload, nsample, cores, steps, warm)
        sample = "%d of the workload's levels per step, %d threads each calling c_ezdefset+c_ezsint on its own levels" % (nsample, cores)
        cb = {"value": v, "unit": unit, "cores": cores, "kind": "reference", "sample": sample,
              "builds": {"oracle_build": {"value": v, "ms_per_step": ms, "flags": "-O2 -ffp-contract=off (the parity oracle's build)"}}}
        if os.path.exists(REF_VEC_SO) and host_can_run_vec_build():
            try:
                v2, ms2, _, _ = cpu_reference_run(args.workload, nsample, cores, max(1, steps // 2), 1, so=REF_VEC_SO)
                cb["builds"]["reference_default_flags"] = {"value": v2, "ms_per_step": ms2,
                                                           "flags": "-O2 -ftree-vectorize -march=x86-64-v3 (gnu.cmake:17-38 with a portable -march)"}
                if v2 > v:   # the faster of the two builds is the baseline the ratio is taken against
                    v, ms = v2, ms2
            except Exception as ex:
                cb["builds"]["reference_default_flags"] = {"value": None, "error": repr(ex)}
        cb["value"] = v
        line = {"impl": "reference", "metric": metric, "value": v, "unit": unit, "n_gpus": args.gpus, "steps": steps, "warmup": warm,
                "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64-on-f32", "data": "synthetic",
                "config": {"workload": desc, "sample": sample},
                "cpu_baseline": cb,
                "e2e": {"value": v, "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return 0

    import torch
    import torch.distributed as dist
    import librmn_b200

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: libezb200.so has no CPU path")
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    B = librmn_b200.load(build_if_missing=True)
    stream = torch.cuda.current_stream()
    B.set_stream(stream.cuda_stream)

    sampler = ClockSampler(local_rank) if rank == 0 else None
    full = not args.no_extras
    m = measure(B, args.workload, args.arith, args.steps, args.warmup, args.levels, rank, world, stream,
                sampler_cb=(sampler.start if sampler else None), sustain_s=(1.0 if full else 0.0), pageable=full, parity=full, gather=args.gather)
    if sampler:
        sampler.stop_flag = True
        sampler.join(timeout=2)

    extra = {}
    if full and world == 1 and args.workload == "c2" and args.arith == 2 and not args.levels:
        for key, (wl, ar) in (("c3_ezuvint", ("c3", 2)), ("c1_ezsint", ("c1", 2)), ("c4_cloud_target", ("c4", 2)), ("c5", ("c5", 2)), ("c2_arith1_fp64", ("c2", 1))):
            try:
                extra[key] = sub_record(measure(B, wl, ar, args.steps, args.warmup, 0, rank, world, stream, sustain_s=0.5, pageable=(wl != "c2"), parity=True))
            except SystemExit as ex:
                extra[key] = {"error": str(ex)}
        B.set_arith(args.arith)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    cpu_baseline = None
    if not args.no_cpu_baseline and os.path.exists(REF_SO):
        try:
            nlev = m["nlev"]
            nsample = min(max(args.cpu_sample_levels, cores), max(nlev, 1))   # at least one level per host thread where the workload has them
            v, ms, _, _ = cpu_reference_run(args.workload, nsample, cores, 2, 1)
            v1, ms1, _, _ = cpu_reference_run(args.workload, 1 if nlev < 4 else 2, 1, 1, 0)   # one thread beside T threads
            cpu_baseline = {"value": v, "unit": unit, "cores": cores, "kind": "reference",
                            "sample": "%d levels of the same workload, %d host threads (one c_ezdefset+c_ezsint stream per thread), %.0f ms/step" % (
                                nsample, cores, ms),
                            "flags": "-O2 -ffp-contract=off (the parity oracle's build)",
                            "single_thread": {"value": v1, "unit": unit, "cores": 1, "sample": "%d level(s), %.0f ms" % (1 if nlev < 4 else 2, ms1)}}
            if os.path.exists(REF_VEC_SO) and host_can_run_vec_build():
                v2, ms2, _, _ = cpu_reference_run(args.workload, nsample, cores, 2, 1, so=REF_VEC_SO)
                cpu_baseline["reference_default_flags"] = {"value": v2, "unit": unit, "cores": cores, "ms_per_step": ms2,
                                                           "flags": "-O2 -ftree-vectorize -march=x86-64-v3 (gnu.cmake:17-38 with a portable -march)"}
        except Exception as ex:   # the baseline is reported, never required for the GPU number
            cpu_baseline = {"value": None, "unit": unit, "cores": cores, "kind": "reference", "sample": "failed: %r" % (ex,)}

    line = {"metric": metric, "value": m["value"], "unit": unit, "n_gpus": world, "steps": args.steps, "warmup": max(3, args.warmup),
            "ms_per_step": m["ms_per_step"], "higher_is_better": True, "scaling": "strong" if m["band_mode"] else "weak", "vs_baseline": None,
            "dtype": "f32" if args.arith == 2 else "f64",
            "data": "synthetic",
            "config": {"workload": m["desc"], "levels_per_gpu": m["nlev"], "n_src": m["nin"], "n_out": m["nout"], "arith": args.arith,
                       "l2": m["l2"],
                       "parallelism": ("output-row bands of one call split over %d GPU(s), replicated gridset and source field" if m["band_mode"] else
                                       "levels partitioned over %d GPU(s), replicated gridset") % world},
            "e2e": m["e2e"],
            "gpu_launches": m["launches"],
            "gridset_build_ms": m["gridset_build_ms"],
            "clocks": sampler.summary() if sampler else None,
            "roofline": m["roofline"],
            "cpu_baseline": cpu_baseline}
    if m["sustained"]:
        line["sustained"] = m["sustained"]
    if m["parity"]:
        line["parity"] = m["parity"]
    if m["collective"]:
        line["collective"] = m["collective"]
    if extra:
        line["extra"] = extra
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
