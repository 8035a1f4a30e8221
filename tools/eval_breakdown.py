"""CUDA-event timing of the fused full-rank eval (lgc_score_topk) against the unfused pair
(lgc_score_block + lgc_topk_rows) on a LightGCN shape (run on the GPU box)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import torch  # noqa: E402
from lgcnhs_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
shape = sys.argv[1] if len(sys.argv) > 1 else "amazon-book"
d = bench.load_shape(shape)
tr, va, _ = d.split()
U, M = d.n_users, d.n_items
torch.manual_seed(42)
xu = torch.empty(U, 64).normal_(std=0.1).to(dev)
xi = torch.empty(M, 64).normal_(std=0.1).to(dev)
seen = ops.seen_csr(torch.from_numpy(d.users[tr]).to(dev), torch.from_numpy(d.items[tr]).to(dev), U, M)


def timeit(fn, n=5):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


from lgcnhs_b200._lib import lib  # noqa: E402

print(f"shape {shape}: U={U} M={M}  score flops = {2.0 * U * M * 64 / 1e9:.1f} GFLOP")
for threads in (256, 512):
    lib().lgc_score_topk_config(threads)
    for k in (20, 100):
        ms = timeit(lambda: ops.score_topk(xu, xi, k, seen, want_values=False))
        print(f"fused score_topk ({threads} thr) k={k:3d}: {ms:8.3f} ms  {U / ms * 1e3:12.0f} users/s  {2.0 * U * M * 64 / ms / 1e9:8.1f} TFLOP/s fp32")
blk = 8192
buf = torch.empty((blk, (M + 3) // 4 * 4), dtype=torch.float32, device=dev)


def unfused(k):
    for u0 in range(0, U, blk):
        u1 = min(u0 + blk, U)
        s = ops.score_block(xu, xi, u0, u1, seen, out=buf[: u1 - u0, :M])
        ops.topk_rows(s, k, want_values=False)


ms = timeit(lambda: unfused(20), 2)
print(f"unfused block + topk k= 20: {ms:8.3f} ms  {U / ms * 1e3:12.0f} users/s")
i1, v1 = ops.score_topk(xu, xi, 20, seen)
s = ops.score_block(xu, xi, 0, 4096, seen)
i2, v2 = ops.topk_rows(s, 20)
print("fused == unfused on the first 4096 users:", bool(torch.equal(i1[:4096], i2) and torch.equal(v1[:4096], v2)))
