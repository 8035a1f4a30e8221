"""Per-row-class cost of the propagation SpMM on one GPU: user rows (gather item embeddings) vs item rows
(gather user embeddings), to size the multi-GPU partition (run on the GPU box)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import torch  # noqa: E402
from lgcnhs_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
d = bench.load_shape(sys.argv[1] if len(sys.argv) > 1 else "ml-20m")
adj_np, _ = bench.train_adj(d)
n = d.n_users + d.n_items
g = ops.NormGraph(torch.from_numpy(adj_np).to(dev), n)
x = (torch.randn(n, 64) * 0.1).to(dev)
out = torch.empty_like(x)
rp = g.rowptr.cpu()


def timeit(a, b, reps=10):
    ch = g.chunk_range(a, b)
    for _ in range(3):
        g.spmm(x, x, 1.0, 1.0, out=out, row_begin=a, row_end=b, chunks=ch)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        g.spmm(x, x, 1.0, 1.0, out=out, row_begin=a, row_end=b, chunks=ch)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3


from lgcnhs_b200._lib import lib  # noqa: E402

for UNR in [int(t) for t in os.environ.get("PROBE_UNROLL", "0").split(",")]:
  lib().lgc_spmm_config(UNR)
  for T in [int(t) for t in os.environ.get("PROBE_LONG_ROW", "0").split(",")]:
    lib().lgc_spmm_long_row(T)
    print("unroll =", UNR, "long_row =", T)
    for name, a, b in (("all rows", 0, n), ("user rows", 0, d.n_users), ("item rows", d.n_users, n),
                       ("1/8 of user rows", 0, d.n_users // 8), ("1/8 of item rows", d.n_users, d.n_users + d.n_items // 8)):
        nnz = int(rp[b] - rp[a])
        us = timeit(a, b)
        print(f"{name:18s} rows {b - a:7d} nnz {nnz:9d}  {us:8.1f} us  {nnz / us / 1e3:7.2f} Gnnz/s")
