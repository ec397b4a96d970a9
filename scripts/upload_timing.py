"""Host -> HBM upload of a numpy float64 array as float32: the library's staged uploader (aisb_upload_as_f32) against
torch's pageable copy + device cast, and its thread scaling. Usage: upload_timing.py [rows] [dims]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from ais_benchmarks_b200 import _device

rows = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
dims = int(sys.argv[2]) if len(sys.argv) > 2 else 8
x = np.random.RandomState(0).rand(rows, dims)
ref = torch.from_numpy(x.astype(np.float32)).cuda()


def wall(fn, reps=5):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter(); out = fn(); torch.cuda.synchronize(); ts.append(time.perf_counter() - t0)
    return min(ts) * 1e3, out


t, out = wall(lambda: torch.from_numpy(x).to("cuda").to(torch.float32))
print("torch pageable float64 copy + device cast : %7.2f ms  (%.1f GB/s of float64)" % (t, x.nbytes / t / 1e6))
assert torch.equal(out, ref)
for th in (1, 2, 3, 4, 6, 8):
    os.environ["AISB_UPLOAD_THREADS"] = str(th)
    t, out = wall(lambda: _device.upload_f32(x))
    assert torch.equal(out, ref), "staged upload differs from numpy astype(float32)"
    print("aisb_upload_as_f32, %d host thread(s)        : %7.2f ms  (%.1f GB/s of float64)" % (th, t, x.nbytes / t / 1e6))
x32 = x.astype(np.float32)
os.environ["AISB_UPLOAD_THREADS"] = "4"
t, out = wall(lambda: _device.upload_f32(x32))
assert torch.equal(out, ref)
t2, _ = wall(lambda: torch.from_numpy(x32).to("cuda"))
print("float32 source: staged 4 threads %.2f ms, torch pageable %.2f ms" % (t, t2))
print("host cores: %d" % (os.cpu_count() or 0))

# the same upload after the process has been idle / blocked on the GPU for as long as a KDE step lasts
import ctypes
y = np.random.RandomState(1).rand(rows, dims)
for label, idle in (("back to back", None), ("after time.sleep(0.37)", "sleep"), ("after a 0.37 s GPU wait", "gpu")):
    for th in (4, 8):
        os.environ["AISB_UPLOAD_THREADS"] = str(th)
        ts = []
        for rep in range(4):
            src = x if rep % 2 == 0 else y
            if idle == "sleep":
                time.sleep(0.37)
            elif idle == "gpu":
                torch.cuda._sleep(int(0.37 * 1.9e9)); torch.cuda.synchronize()
            t0 = time.perf_counter(); out = _device.upload_f32(src); torch.cuda.synchronize(); ts.append((time.perf_counter() - t0) * 1e3)
        print("%-26s %d threads: %s ms" % (label, th, " ".join("%.2f" % t for t in ts)))
