"""CPU simulation (numpy): what fraction of (warp of points, chunk of centres) units of the configs[3] KDE sweep could
be skipped WITHOUT evaluating any pair, from axis-aligned bounding boxes of Z-ordered groups alone.
Result at full size (24 groups): 45.7 % of the 128 x 128 units hold a pair within the margin, only 4.3 % can be
excluded from the boxes -- in 8 dimensions the boxes of 128 neighbours overlap almost everywhere, so the sweep evaluates
every pair and skips only the exponentials.
Usage: kde_bbox_sim.py [scale] [point groups sampled] [points per group] [centres per chunk] [morton bits]"""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import kde_workload, KDE_CFG


def morton(x, bits=4):
    lo = x.min(0); span = np.maximum(x.max(0) - lo, 1e-30)
    q = np.clip(((x - lo) / span * (1 << bits)).astype(np.int64), 0, (1 << bits) - 1)
    code = np.zeros(len(x), dtype=np.int64)
    for d in range(x.shape[1]):
        for b in range(bits):
            code |= ((q[:, d] >> b) & 1) << (b * x.shape[1] + d)
    return np.argsort(code, kind="stable")


scale = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
ngroups = int(sys.argv[2]) if len(sys.argv) > 2 else 64
PG = int(sys.argv[3]) if len(sys.argv) > 3 else 128
CG = int(sys.argv[4]) if len(sys.argv) > 4 else 128
bits = int(sys.argv[5]) if len(sys.argv) > 5 else 4
mu, sig, w, sup, centres, x = kde_workload(scale=scale)
a = np.sqrt(np.log2(np.e) / (2 * KDE_CFG["bw"]))
cs = centres[morton(centres, bits)].astype(np.float64) * a
xs = x[morton(x, bits)].astype(np.float64) * a
K = len(cs) // CG * CG; n = len(xs)
cs = cs[:K]
marg = 21 + np.log2(K)
clo = cs.reshape(-1, CG, cs.shape[1]).min(1); chi = cs.reshape(-1, CG, cs.shape[1]).max(1)
v2 = (cs ** 2).sum(1)
rs = np.random.RandomState(3)
tot = skip_ideal = 0
skip_centroid = 0
need_units = 0
for g in range(ngroups):
    g0 = rs.randint(n // PG) * PG
    u = xs[g0:g0 + PG]; u2 = (u ** 2).sum(1)
    dmin = np.full(PG, np.inf)
    unit_min = np.full((PG, K // CG), np.inf)
    for c0 in range(0, K, 65536):
        c1 = min(K, c0 + 65536)
        d2 = u2[:, None] + v2[None, c0:c1] - 2 * u @ cs[c0:c1].T
        dmin = np.minimum(dmin, d2.min(1))
        unit_min[:, c0 // CG:c1 // CG] = d2.reshape(PG, -1, CG).min(2)
    plo = u.min(0); phi = u.max(0)
    gap = np.maximum(0, np.maximum(clo - phi, plo - chi))       # [chunks, D]
    lb = (gap ** 2).sum(1)
    thr = dmin.max() + marg
    tot += len(lb); skip_ideal += (lb > thr).sum()
    need = (unit_min < (dmin[:, None] + marg)).any(0)             # unit has a pair within the margin for some point
    need_units += need.sum()
print("scale %g K=%d PG=%d CG=%d bits=%d: units %d, needed %.4f, bbox-skippable (final minima known) %.4f"
      % (scale, K, PG, CG, bits, tot, need_units / tot, skip_ideal / tot))
