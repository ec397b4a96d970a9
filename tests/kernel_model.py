"""numpy model of what the CUDA kernels compute FROM THE PACKED BUFFERS (include/ais_b200.h layout):
log2-domain t_k = c_k - ||A_k x + b_k||^2, then LSE. Lets the CPU suite validate the float64->packed
conversion (aisb_pack_mixture_f64) and the layout without a GPU."""
import numpy as np

LN2 = np.log(2.0)


def packed_logpdf(x, rows, offs, K, D, kind, iso_scale, DP):
    x = np.asarray(x, dtype=np.float64)
    n = x.shape[0]
    t = np.empty((n, K))
    rows = rows.astype(np.float64)
    for k in range(K):
        row = rows[k]
        if kind == 0:
            z = x * float(iso_scale) + row[:D]
        elif kind == 1:
            z = x * row[:D] + row[DP:DP + D]
        else:
            tri = DP * (DP + 1) // 2
            z = np.empty((n, D))
            for r in range(D):
                a = row[r * (r + 1) // 2: r * (r + 1) // 2 + r + 1]
                z[:, r] = x[:, :r + 1] @ a + row[tri + r]
        t[:, k] = float(offs[k]) - (z * z).sum(1)
    m = t.max(1)
    with np.errstate(divide="ignore", invalid="ignore"):
        s = np.exp2(t - np.where(np.isfinite(m), m, 0.0)[:, None]).sum(1)
        return (np.where(np.isfinite(m), m, 0.0) + np.log2(s)) * LN2
