"""Experiment: L1 policy per source row in the propagation SpMM (ml-20m train graph).  colidx gets a flag in bit 31 for the
H highest-degree nodes; flagged rows are gathered with L1::evict_last, the rest with L1::no_allocate."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import torch  # noqa: E402
from lgcnhs_b200 import ops  # noqa: E402
from lgcnhs_b200._lib import lib  # noqa: E402

dev = torch.device("cuda:0")
shape = sys.argv[1] if len(sys.argv) > 1 else "ml-20m"
d = bench.load_shape(shape)
adj_np, _ = bench.train_adj(d)
n = d.n_users + d.n_items
g = ops.NormGraph(torch.from_numpy(adj_np).to(dev), n)
torch.manual_seed(42)
x0 = (torch.randn(n, 64) * 0.1).to(dev)
out = torch.empty_like(x0)
ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]


def timed(reps=10):
    for _ in range(3):
        g.spmm(x0, x0, 1.0, 1.0, out=out)
    torch.cuda.synchronize()
    ev[0].record()
    for _ in range(reps):
        g.spmm(x0, x0, 1.0, 1.0, out=out)
    ev[1].record()
    torch.cuda.synchronize()
    return ev[0].elapsed_time(ev[1]) / reps * 1e3


base = timed()
ref = out.clone()
print(f"{shape}: plain {base:.1f} us/layer")
deg = (g.rowptr[1:] - g.rowptr[:-1]).to(torch.int64)
plain_col = g.colidx.clone()
for H in (0, 256, 512, 768, 1024, 2048, 4096):
    hot = torch.zeros(n, dtype=torch.bool, device=dev)
    if H:
        hot[torch.topk(deg, H).indices] = True
    cov = float(deg[hot].sum()) / g.nnz
    flagged = plain_col.to(torch.int64)
    flagged = torch.where(hot[flagged], flagged | (1 << 31), flagged)
    g.colidx = (flagged & 0xFFFFFFFF).to(torch.int64).sub_(torch.where(flagged >= (1 << 31), 1 << 32, 0)).to(torch.int32)
    lib().lgc_spmm_hot_mode(1)
    t = timed()
    same = bool(torch.equal(out, ref))
    lib().lgc_spmm_hot_mode(0)
    g.colidx = plain_col
    print(f"  hot set {H:5d} rows ({H * 256 / 1024:.0f} KB), covers {100 * cov:5.1f}% of the gathers: {t:.1f} us/layer "
          f"({base / t:.3f}x), bit-identical={same}")
