"""distCUDA2 timing: ours vs the compiled reference simple-knn (oracle/_ref/libref_knn.so) on the same GPU."""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from semantic_gaussians_b200.simple_knn import distCUDA2
dev = torch.device("cuda:0")
lib = C.CDLL(os.path.join(ROOT, "oracle", "_ref", "libref_knn.so")); lib.ref_knn.argtypes = [C.c_int, C.c_void_p, C.c_void_p]
rng = np.random.default_rng(0)
for n, kind in ((100000, "uniform"), (1000000, "uniform"), (1000000, "room")):
    pts = rng.uniform(-1.3, 1.3, (n, 3)) if kind == "uniform" else np.concatenate(
        [rng.uniform(-4, 4, (n, 2)), rng.uniform(-1.5, 1.5, (n, 1)) * (rng.random((n, 1)) < 0.2) + 1.5 * np.sign(rng.standard_normal((n, 1))) * (rng.random((n, 1)) > 0.5)], axis=1)
    p = torch.from_numpy(pts.astype(np.float32)).to(dev)
    out = torch.zeros(n, device=dev)
    def t(fn, reps=3):
        fn(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps): fn()
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps
    ours = t(lambda: distCUDA2(p)); ref = t(lambda: lib.ref_knn(n, p.data_ptr(), out.data_ptr()))
    print(f"{kind} P={n}: ours {ours:.2f} ms | reference simple-knn {ref:.2f} ms | equal {torch.equal(distCUDA2(p), out)}", flush=True)
