"""CPU simulation (numpy) of how many (warp, centre-group) units of the configs[3] KDE sweep are NOT negligible,
for different skip granularities. Used to choose the screen/refine granularity of the D = 8 pairwise kernel."""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import kde_workload, KDE_CFG

def morton(x, bits=4):
    lo = x.min(0); span = np.maximum(x.max(0) - lo, 1e-30)
    q = np.clip(((x - lo) / span * (1 << bits)).astype(np.int64), 0, (1 << bits) - 1)
    D = x.shape[1]
    code = np.zeros(len(x), dtype=np.int64)
    for d in range(D):
        for b in range(bits):
            code |= ((q[:, d] >> b) & 1) << (b * D + d)
    return np.argsort(code, kind="stable")

scale = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
nwarps = int(sys.argv[2]) if len(sys.argv) > 2 else 24
nsplit = int(sys.argv[3]) if len(sys.argv) > 3 else 1
margin_base = float(sys.argv[4]) if len(sys.argv) > 4 else 28.0
seeded = int(sys.argv[5]) if len(sys.argv) > 5 else 0      # 1: running max starts at the final max (ideal seeding)
layout = sys.argv[6] if len(sys.argv) > 6 else "strided"   # strided: e*128+32w+l ; consec: warp owns 128 consecutive points
mu, sig, w, sup, centres, x = kde_workload(scale=scale)
a = np.sqrt(np.log2(np.e) / (2 * KDE_CFG["bw"]))
cs = centres[morton(centres)].astype(np.float64) * a
xs = x[morton(x)].astype(np.float64) * a
K = len(cs); n = len(xs)
margin = margin_base + np.log2(K)
rs = np.random.RandomState(5)
ntile = n // 512
res = {}
def add(k, num, den):
    a_, b_ = res.get(k, (0, 0)); res[k] = (a_ + num, b_ + den)
v2 = (cs ** 2).sum(1)
for it in range(nwarps):
    tile = rs.randint(ntile); wi = rs.randint(4)
    if layout == "strided":
        idx = np.array([tile * 512 + e * 128 + 32 * wi + l for e in range(4) for l in range(32)])  # [e*32+l]
    else:
        idx = np.array([tile * 512 + 128 * wi + e * 32 + l for e in range(4) for l in range(32)])
    u = xs[idx]; u2 = (u ** 2).sum(1)
    per = (K + nsplit - 1) // nsplit
    for s in range(nsplit):
        m = np.full(128, -3e38)
        if seeded:
            for c0 in range(0, K, 65536):
                m = np.maximum(m, (-(u2[:, None] + v2[None, c0:c0 + 65536] - 2 * u @ cs[c0:c0 + 65536].T)).max(1))
        for c0 in range(s * per, min(K, (s + 1) * per), 65536):
            c1 = min(K, (s + 1) * per, c0 + 65536)
            C = (c1 - c0) // 8 * 8
            if C == 0: continue
            t = -(u2[:, None] + v2[None, c0:c0 + C] - 2 * u @ cs[c0:c0 + C].T)   # [128, C]
            G = C // 8
            gm = t.reshape(128, G, 8).max(2)                                      # [128, G]
            cm = np.maximum.accumulate(np.concatenate([m[:, None], gm], 1), 1)    # running max before group g = cm[:, g]
            before = cm[:, :G]
            need_pair = t.reshape(128, G, 8) > (before[:, :, None] - margin)      # [128, G, 8]
            add("pair", need_pair.sum(), need_pair.size)
            np8 = need_pair.any(2)                                                # point x group8
            add("point x g8", np8.sum(), np8.size)
            # warp-level units: rows are [e*32 + lane]
            r = need_pair.reshape(4, 32, G, 8)
            add("32pts x c1", r.any(1).sum(), 4 * G * 8)
            add("64pts(pp) x c1", r.reshape(2, 64, G, 8).any(1).sum(), 2 * G * 8)
            add("128pts x c1", r.reshape(128, G, 8).any(0).sum(), G * 8)
            add("32pts x g4", r.reshape(4, 32, G * 2, 4).any(3).any(1).sum(), 4 * G * 2)
            add("64pts(pp) x g4", r.reshape(2, 64, G * 2, 4).any(3).any(1).sum(), 2 * G * 2)
            add("64pts(pp) x g8", r.reshape(2, 64, G, 8).any(3).any(1).sum(), 2 * G)
            add("128pts x g8", r.reshape(128, G, 8).any(2).any(0).sum(), G)
            add("128pts x g4", r.reshape(128, G * 2, 4).any(2).any(0).sum(), G * 2)
            # max over lanes of per-lane needed pair count per (128pts x g8) unit: divergent per-lane refine cost
            cnt = need_pair.reshape(4, 32, G, 8).sum(3).sum(0)                    # [32 lanes, G] pairs per lane per group
            add("lane-divergent iterations per g8 (max over lanes)", cnt.max(0).sum(), G)
            add("lane mean per g8", cnt.mean(0).sum(), G)
            cnt_e = need_pair.reshape(4, 32, G, 8).sum(3)                         # [e, lane, G]
            add("sum_e max_lanes popc(mask_e) per g8", cnt_e.max(1).sum(), G)
            add("warp-units (128pts x g8) with any needed pair", (cnt.max(0) > 0).sum(), G)
            m = cm[:, G]
print("scale %g K=%d n=%d nsplit=%d margin=%.1f warps=%d seeded=%d layout=%s" % (scale, K, n, nsplit, margin, nwarps, seeded, layout))
for k, (a_, b_) in res.items():
    print("%-55s %.4f" % (k, a_ / b_))
