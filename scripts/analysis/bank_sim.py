"""CPU-side model of shared-memory bank conflicts of the window tile kernel on C2's Yang tiles.
For each 32x32 target tile: window base (min c0, min r0 over the 2x2 quad) per thread; for a thread->quad mapping and a
staging pitch, wavefronts per LDS = max over banks of the number of DISTINCT addresses hitting the bank."""
import numpy as np, sys, itertools

NI, NJ = 3600, 1801
lon = (np.arange(NI) * 0.1)[None, :].repeat(NJ, 0)
lat = (-90 + np.arange(NJ) * 0.1)[:, None].repeat(NI, 1)
la, lo = np.radians(lat), np.radians(lon)
x, y, z = np.cos(la) * np.cos(lo), np.cos(la) * np.sin(lo), np.sin(la)
yin = (lon >= 45) & (lon <= 315) & (np.abs(lat) <= 45)
xr, yr, zr = -x, -z, -y
latr = np.degrees(np.arcsin(np.clip(zr, -1, 1)))
lonr = np.degrees(np.arctan2(yr, xr)) % 360
px = (lonr - 42) / 0.15 + 1
py = (latr + 48) / 0.15 + 1
c0 = np.floor(px).astype(np.int64) - 2
r0 = np.floor(py).astype(np.int64) - 2
SRC_NI = 1841

def tiles():
    for by in range((NJ + 31) // 32):
        for bx in range((NI + 31) // 32):
            j0, i0 = by * 32, bx * 32
            if j0 + 32 > NJ or i0 + 32 > NI: continue
            if yin[j0:j0 + 32, i0:i0 + 32].any(): continue
            yield bx, by

def quad_bases(bx, by):
    j0, i0 = by * 32, bx * 32
    C = c0[j0:j0 + 32, i0:i0 + 32].reshape(16, 2, 16, 2)
    R = r0[j0:j0 + 32, i0:i0 + 32].reshape(16, 2, 16, 2)
    wc = C.min(axis=(1, 3)); wr = R.min(axis=(1, 3))   # [qy][qx]
    spanc = C.max(axis=(1, 3)) - wc; spanr = R.max(axis=(1, 3)) - wr
    return wc, wr, spanc, spanr

def wavefronts(addr):      # addr: [nwarp][32] word addresses
    out = []
    for w in addr:
        u = np.unique(w)
        out.append(np.bincount(u % 32, minlength=32).max())
    return np.array(out)

def wavefronts64(addr):    # LDS.64: half-warps, 8-byte units, 16 banks of 8 B... model: addr in words (even), per half warp
    out = []
    for w in addr:
        t = 0
        for h in (w[:16], w[16:]):
            u = np.unique(h // 2)
            t += np.bincount(u % 16, minlength=16).max()
        out.append(t)
    return np.array(out)

shapes = {"16x2": (16, 2), "8x4": (8, 4), "4x8": (4, 8), "32x1": (32, 1), "2x16": (2, 16)}
def warp_lanes(shape):
    sx, sy = shapes[shape]
    warps = []
    for wy in range(16 // sy if sy <= 16 else 1):
        for wx in range(16 // sx if sx <= 16 else 1):
            lanes = [(wy * sy + ly, wx * sx + lx) for ly in range(sy) for lx in range(sx)]
            warps.append(lanes)
    if shape == "32x1":
        warps = [[(q, lx) for lx in range(16)] + [(q + 1, lx) for lx in range(16)] for q in range(0, 16, 2)]
        # 32x1 impossible with 16 quads per row; same as 16x2
    return warps

rng = np.random.default_rng(0)
T = list(tiles())
print("yang-only tiles:", len(T))
sel = [T[i] for i in rng.choice(len(T), 400, replace=False)]
res = {}
for bx, by in sel:
    wc, wr, sc, sr = quad_bases(bx, by)
    w = int((wc.max() + 4) - wc.min() + 1)
    base_pitch = (w + 3) + (((SRC_NI - (w + 3)) % 4) + 4) % 4
    for shape in ("16x2", "8x4", "4x8", "2x16"):
        lanes = warp_lanes(shape)
        for dp in range(0, 32, 4):
            pitch = base_pitch + dp
            addr = np.array([[(wr[qy, qx] - wr.min()) * pitch + (wc[qy, qx] - wc.min()) for qy, qx in L] for L in lanes])
            wf = wavefronts(addr).mean()
            res.setdefault((shape, dp), []).append(wf)
            if dp == 0:
                res.setdefault((shape, "lds64"), []).append(wavefronts64(addr & ~1).mean())
    # best per tile
import collections
print("mean LDS.32 wavefronts per warp-LDS, fixed (shape, extra pitch):")
for k in sorted(res, key=lambda k: (k[0], str(k[1]))):
    print(k, "%.3f" % np.mean(res[k]))
allk = [k for k in res if k[1] != "lds64"]
M = np.array([res[k] for k in allk])
print("best-per-tile over all (shape,pitch): %.3f" % M.min(axis=0).mean())
for shape in shapes:
    ks = [k for k in allk if k[0] == shape]
    if ks: print("best-per-tile pitch for shape", shape, "%.3f" % np.array([res[k] for k in ks]).min(axis=0).mean())

# ---- pairing of quad rows (each half-warp = one row of 16 quads, the two halves of a warp = any two quad rows) ----
def mult(addrs):
    u = np.unique(addrs)
    return np.bincount(u % 32, minlength=32).max()
tot = {"adjacent": [], "bestpair": [], "bestpair_pitch": [], "intra": []}
for bx, by in sel:
    wc, wr, sc, sr = quad_bases(bx, by)
    w = int((wc.max() + 4) - wc.min() + 1)
    base_pitch = (w + 3) + (((SRC_NI - (w + 3)) % 4) + 4) % 4
    best_over_pitch = 1e9
    for dp in range(0, 32, 4):
        pitch = base_pitch + dp
        A = (wr - wr.min()) * pitch + (wc - wc.min())     # [16 qy][16 qx]
        cost = np.zeros((16, 16))
        for a in range(16):
            for b in range(a + 1, 16):
                cost[a, b] = cost[b, a] = mult(np.concatenate([A[a], A[b]]))
        # greedy matching + 2-opt
        left = list(range(16)); pairs = []
        while left:
            a = left.pop(0)
            b = min(left, key=lambda q: cost[a, q]); left.remove(b); pairs.append((a, b))
        improved = True
        while improved:
            improved = False
            for i in range(8):
                for j in range(i + 1, 8):
                    (a, b), (c, d) = pairs[i], pairs[j]
                    cur = cost[a, b] + cost[c, d]
                    for (p, q, r, s) in ((a, c, b, d), (a, d, b, c)):
                        if cost[p, q] + cost[r, s] < cur:
                            pairs[i], pairs[j] = (p, q), (r, s); cur = cost[p, q] + cost[r, s]; improved = True
        m = np.mean([cost[a, b] for a, b in pairs])
        if dp == 0:
            tot["adjacent"].append(np.mean([cost[2 * k, 2 * k + 1] for k in range(8)]))
            tot["bestpair"].append(m)
            tot["intra"].append(np.mean([mult(A[a]) for a in range(16)]))
        best_over_pitch = min(best_over_pitch, m)
    tot["bestpair_pitch"].append(best_over_pitch)
for k, v in tot.items(): print(k, "%.3f" % np.mean(v))
