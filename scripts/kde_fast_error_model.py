"""Float32 error model of the KDE fast kernel (csrc/pairwise.cuh, mixture_pass_fast), measured on the CPU with numpy.

Emulates, with exact float32 rounding (every FMA = float64 product + sum rounded once), three ways of evaluating the pair
exponent t = -||a x + b_j||^2 on warps of 128 Z-ordered evaluation points of the configs[3] workload:
  centred    dl = fma(x, a, b); t -= dl*dl                               (the exact kernels)
  expanded   t = -||a x||^2 - 2a (h_j + x . b_j)                          (no shift)
  recentred  the same after shifting points and rows by the warp centroid (the fast kernel)
and prints, per warp, the worst |log p - float64 reference| / max(1, |log p|) and the ratio of the recentred error to
2^-24 (||a x'||^2 + |t_max| + 1) -- the predictor behind kFastErrC in csrc/mixture.cu (observed maximum 2.4 over 5 120
points at scale 0.2; the kernel uses 4 and re-evaluates points whose estimate exceeds 3e-6 max(1, |log p|) exactly; on the
GPU, 1e6 points at full size: max deviation from the centred kernel 2.0e-6 with, 5.8e-6 without the re-evaluation).
Usage: kde_fast_error_model.py [scale] [warps]"""
import sys, os, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import kde_workload, KDE_CFG
f32=np.float32
def morton(x, bits=4):
    lo = x.min(0); span = np.maximum(x.max(0) - lo, 1e-30)
    q = np.clip(((x - lo) / span * (1 << bits)).astype(np.int64), 0, (1 << bits) - 1)
    D = x.shape[1]; code = np.zeros(len(x), dtype=np.int64)
    for d in range(D):
        for b in range(bits): code |= ((q[:, d] >> b) & 1) << (b * D + d)
    return np.argsort(code, kind="stable")
scale=float(sys.argv[1]) if len(sys.argv)>1 else 0.05; nw=int(sys.argv[2]) if len(sys.argv)>2 else 4
mu, sig, w, sup, centres, x = kde_workload(scale=scale)
a64 = np.sqrt(np.log2(np.e)/(2*KDE_CFG["bw"])); a=f32(a64)
cs = centres[morton(centres)]; xs = x[morton(x)]
K=len(cs); n=len(xs)
b = (-(a*cs)).astype(f32)        # packed rows: f32(-a*mu) (kernel computes -a*samples in f32)
rs=np.random.RandomState(1)
allr=[]
worst={"centred":0,"expanded":0,"recentred":0}
def fma(x,y,z): return (x.astype(np.float64)*y.astype(np.float64)+z.astype(np.float64)).astype(f32)
for it in range(nw):
    i0 = rs.randint(n//128)*128
    X = xs[i0:i0+128]                      # f32 [128,8]
    # exact reference (f64) from the f32 inputs
    d2 = ((X.astype(np.float64)[:,None,:]-cs.astype(np.float64)[None,:,:])**2).sum(-1) if K<=20000 else None
    t_ref = np.empty((128,K))
    for c0 in range(0,K,50000):
        C=cs[c0:c0+50000].astype(np.float64)
        t_ref[:,c0:c0+50000] = -(a64**2)*((X.astype(np.float64)[:,None,:]-C[None])**2).sum(-1)
    M = t_ref.max(1,keepdims=True)
    lp_ref = (M[:,0]+np.log2(np.exp2(t_ref-M).sum(1)))*np.log(2) - np.log(K) - 4*np.log(2*np.pi*0.02)
    def lp_from(t):
        t=t.astype(np.float64); Mx=t.max(1,keepdims=True)
        return (Mx[:,0]+np.log2(np.exp2(t-Mx).sum(1)))*np.log(2) - np.log(K) - 4*np.log(2*np.pi*0.02)
    # (1) centred f32: dl=fma(x,a,b); acc=fma(-dl,dl,acc)
    acc=np.zeros((128,K),f32)
    for d in range(8):
        dl = fma(X[:,d:d+1], np.full((1,1),a,f32), b[None,:,d])
        acc = fma(-dl, dl, acc)
    e1 = np.abs(lp_from(acc)-lp_ref)/np.maximum(1,np.abs(lp_ref))
    # (2) expanded, uncentred: s = h + x.b ; t = -(u2 + 2a s)
    inv2a=f32(0.5)/a
    b2=np.zeros(K,f32)
    for d in range(8): b2=fma(b[:,d],b[:,d],b2)
    h=(b2*inv2a).astype(f32)
    u2=np.zeros(128,f32)
    for d in range(8): u2=fma(X[:,d],X[:,d],u2)
    u2=(u2*(a*a)).astype(f32)
    s=np.broadcast_to(h[None,:],(128,K)).copy()
    for d in range(8): s=fma(X[:,d:d+1], b[None,:,d], s)
    t2=fma(s, np.full((1,1),f32(-2)*a,f32), -u2[:,None])
    e2 = np.abs(lp_from(t2)-lp_ref)/np.maximum(1,np.abs(lp_ref))
    # (3) recentred: c = mean of the warp's points; x' = x - c ; b' = b + a*c
    c=X.astype(np.float64).mean(0).astype(f32)
    Xp=(X-c[None,:]).astype(f32)
    ac=(a*c).astype(f32)
    bp=(b+ac[None,:]).astype(f32)
    b2=np.zeros(K,f32)
    for d in range(8): b2=fma(bp[:,d],bp[:,d],b2)
    h=(b2*inv2a).astype(f32)
    u2=np.zeros(128,f32)
    for d in range(8): u2=fma(Xp[:,d],Xp[:,d],u2)
    u2=(u2*(a*a)).astype(f32)
    s=np.broadcast_to(h[None,:],(128,K)).copy()
    for d in range(8): s=fma(Xp[:,d:d+1], bp[None,:,d], s)
    t3=fma(s, np.full((1,1),f32(-2)*a,f32), -u2[:,None])
    e3 = np.abs(lp_from(t3)-lp_ref)/np.maximum(1,np.abs(lp_ref))
    for k,e in (("centred",e1),("expanded",e2),("recentred",e3)): worst[k]=max(worst[k],e.max())
    ab3=np.abs(lp_from(t3)-lp_ref); ab2=np.abs(lp_from(t2)-lp_ref); ab1=np.abs(lp_from(acc)-lp_ref)
    Uw=float(u2.max())
    tm = -t_ref.max(1)
    ratio = ab3/(2.0**-24*(u2.astype(np.float64)+tm+1.0))
    ratio2 = ab3/(2.0**-24*(u2.astype(np.float64)+1.0))
    allr.append(np.stack([ab3,u2.astype(np.float64),tm,lp_ref],1))
    print("   err/(2^-24 (U_i+|tmax|+1)) max %.3f  ; err/(2^-24(U_i+1)) max %.3f"%(ratio.max(),ratio2.max()))
    print("   U_w=%.1f  abs err: centred %.2e expanded %.2e recentred %.2e ; at argmax recentred: lp=%.2f"%(Uw,ab1.max(),ab2.max(),ab3.max(),lp_ref[ab3.argmax()]))
    print(it, "lp range %.1f..%.1f"%(lp_ref.min(),lp_ref.max()), "centred %.2e expanded %.2e recentred %.2e"%(e1.max(),e2.max(),e3.max()), flush=True)
print(worst)
A=np.concatenate(allr); err,U,tm,lp=A.T
r=err/(2.0**-24*(U+tm+1)); print('err / (2^-24 (U + |t_max| + 1)): max %.3f  p99 %.3f  median %.3f over %d points'%(r.max(),np.percentile(r,99),np.median(r),len(r)))
