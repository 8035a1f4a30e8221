"""K-layer propagation on small graphs: per-layer launches vs the one-launch cooperative kernel, eager and replayed from a
CUDA graph, for 1/2/4 resident CTAs per SM.  Device time per K=3 call."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import torch  # noqa: E402
from lgcnhs_b200 import ops  # noqa: E402
from lgcnhs_b200._lib import lib  # noqa: E402

dev = torch.device("cuda:0")
ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]


def timed(fn, reps=50):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    ev[0].record()
    for _ in range(reps):
        fn()
    ev[1].record()
    torch.cuda.synchronize()
    return ev[0].elapsed_time(ev[1]) / reps * 1e3


def graphed(fn):
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        fn()
    return g.replay


for shape in sys.argv[1:] or ["ml-100k", "douban", "ml-1m"]:
    d = bench.load_shape(shape)
    adj_np, _ = bench.train_adj(d)
    n = d.n_users + d.n_items
    g = ops.NormGraph(torch.from_numpy(adj_np).to(dev), n)
    torch.manual_seed(0)
    x0 = (torch.randn(n, 64) * 0.1).to(dev)
    out = torch.empty_like(x0)
    tmp = (torch.empty_like(x0), torch.empty_like(x0))
    plain = lambda: g.propagate_mean(x0, 3, out=out, tmp=tmp, coop=False)  # noqa: E731
    coop = lambda: g.propagate_mean(x0, 3, out=out, tmp=tmp, coop=True)  # noqa: E731
    cu = g.coop_units()
    line = f"{shape:9s} nnz={g.nnz:8d} units={cu['n_units']:6d} split rows={cu['n_split']:4d} | per-layer launches: eager {timed(plain):6.1f} us, graph {timed(graphed(plain)):6.1f} us"
    for c in (1, 2, 4):
        lib().lgc_coop_config(c)
        line += f" | coop x{c}: eager {timed(coop):6.1f} us, graph {timed(graphed(coop)):6.1f} us"
    lib().lgc_coop_config(2)
    print(line, flush=True)
