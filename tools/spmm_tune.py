"""Time the propagation for the SpMM unroll variants on several shapes (run on the GPU box)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import torch  # noqa: E402
from lgcnhs_b200 import ops  # noqa: E402
from lgcnhs_b200._lib import lib, check  # noqa: E402

dev = torch.device("cuda:0")
shapes = sys.argv[1:] or ["ml-1m", "amazon-book", "ml-20m"]
for shape in shapes:
    d = bench.load_shape(shape)
    adj_np, _ = bench.train_adj(d)
    n = d.n_users + d.n_items
    g = ops.NormGraph(torch.from_numpy(adj_np).to(dev), n)
    x0 = (torch.randn(n, 64) * 0.1).to(dev)
    E = torch.empty_like(x0); tmp = (torch.empty_like(x0), torch.empty_like(x0))
    ref = None
    for un in (2, 4, 8):
        check(lib().lgc_spmm_config(un))
        for _ in range(5):
            g.propagate_mean(x0, 3, out=E, tmp=tmp)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            g.propagate_mean(x0, 3, out=E, tmp=tmp)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 20
        same = True if ref is None else torch.equal(ref, E)
        ref = E.clone() if ref is None else ref
        gbs = bench.prop_bytes(g.nnz, n) / ms / 1e6
        print(f"{shape:12s} nnz={g.nnz:9d} unroll={un}: {ms*1e3/3:8.1f} us/layer  {gbs:8.0f} GB/s algorithmic  bit-identical={same}", flush=True)
