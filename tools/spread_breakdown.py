"""Per-stage CUDA-event timing of one spreading lambda-step (run on the GPU box)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import numpy as np, torch  # noqa: E402
from lgcnhs_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
shape = sys.argv[1] if len(sys.argv) > 1 else "ml-1m"
d = bench.load_shape(shape)
tr, va, _ = d.split()
sel = np.concatenate([tr, va])
w_mode = sys.argv[2] if len(sys.argv) > 2 else "u8x4"
eng = ops.SpreadingEngine(d.n_users, d.n_items, torch.from_numpy(d.users[sel]).to(dev), torch.from_numpy(d.items[sel]).to(dev), w_mode=w_mode)
U, M = d.n_users, d.n_items
eng.general_w(); eng.scale(0.3)
F = torch.empty((U, (M + 3) // 4 * 4), dtype=torch.float32, device=dev)[:, :M]
excl = eng.excl

def timeit(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3

print(f"shape {shape}: U={U} M={M} w_mode={w_mode}")
print("general_w (pack + u8 gemm) %.1f us" % timeit(lambda: eng.general_w(), 5))
print("scale_w                    %.1f us" % timeit(lambda: eng.scale(0.3)))
print("resource gemm              %.1f us" % timeit(lambda: eng.resource(out=F)))
for k in (10, 20, 100):
    print("topk k=%3d filtered        %.1f us" % (k, timeit(lambda: ops.topk_rows(F, k, excl))))
print("topk k= 20 unfiltered      %.1f us" % timeit(lambda: ops.topk_rows(F, 20, None)))
print("topk k= 20 no values       %.1f us" % timeit(lambda: ops.topk_rows(F, 20, excl, want_values=False)))
print("recommend(lam) total       %.1f us" % timeit(lambda: eng.recommend(0.3, 20, F_out=F)))
