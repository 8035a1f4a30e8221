"""Short, deterministic launch sequence for ncu (never a bench value): a few propagation steps
on the ml-20m train graph and a few spreading lambda-steps on the ml-1m shape."""
import argparse, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402  (sets sys.path for the package)
import numpy as np, torch  # noqa: E402
from lgcnhs_b200 import ops  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--what", default="prop", choices=["prop", "spread", "both", "eval", "fusedtopk", "all"])
ap.add_argument("--shape", default="ml-20m")
ap.add_argument("--steps", type=int, default=2)
a = ap.parse_args()
dev = torch.device("cuda:0")
if a.what in ("prop", "both", "all"):
    d = bench.load_shape(a.shape)
    adj_np, _ = bench.train_adj(d)
    n = d.n_users + d.n_items
    g = ops.NormGraph(torch.from_numpy(adj_np).to(dev), n)
    torch.manual_seed(42)
    x0 = (torch.randn(n, 64) * 0.1).to(dev)
    for _ in range(a.steps + 1):
        E = g.propagate_mean(x0, 3)
    torch.cuda.synchronize()
    print("prop ok", float(E.abs().sum()))
if a.what in ("spread", "both", "all"):
    d = bench.load_shape("ml-1m")
    tr, va, _ = d.split()
    sel = np.concatenate([tr, va])
    eng = ops.SpreadingEngine(d.n_users, d.n_items, torch.from_numpy(d.users[sel]).to(dev), torch.from_numpy(d.items[sel]).to(dev))
    for lam in np.linspace(0.2, 0.8, a.steps + 1):
        idx, val = eng.recommend(float(lam), 20)
    torch.cuda.synchronize()
    print("spread ok", int(idx.sum()))
if a.what in ("eval", "all"):
    # fused full-rank top-20 (lgc_score_topk) on the amazon-book shape
    d = bench.load_shape("amazon-book")
    tr, va, _ = d.split()
    torch.manual_seed(42)
    xu = (torch.randn(d.n_users, 64) * 0.1).to(dev)
    xi = (torch.randn(d.n_items, 64) * 0.1).to(dev)
    seen = ops.seen_csr(torch.from_numpy(d.users[tr]).to(dev), torch.from_numpy(d.items[tr]).to(dev), d.n_users, d.n_items)
    for _ in range(2):
        idx, _ = ops.score_topk(xu, xi, 20, seen, want_values=False)
    torch.cuda.synchronize()
    print("eval ok", int(idx.sum()))
if a.what in ("fusedtopk", "all"):
    # hs_resource_topk: top-20 selected inside the F-GEMM epilogue (ml-1m shape), and the symmetric G schedule
    d = bench.load_shape("ml-1m")
    tr, va, _ = d.split()
    sel = np.concatenate([tr, va])
    eng = ops.SpreadingEngine(d.n_users, d.n_items, torch.from_numpy(d.users[sel]).to(dev), torch.from_numpy(d.items[sel]).to(dev))
    eng.general_w()
    for lam in (0.3, 0.6):
        idx, val = eng.recommend(lam, 20, fused=True)
    torch.cuda.synchronize()
    print("fusedtopk ok", int(idx.sum()))
