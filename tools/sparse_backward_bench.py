"""CUDA-event timing of the fused BPR training step with the masked (sparse-source) first gradient layer against the
dense backward (LGCNHS_DENSE_BACKWARD=1), plus the two variants of that layer alone (GPU box)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench  # noqa: E402
import _stub_const  # noqa: E402
import torch  # noqa: E402

_stub_const.install()
from lgcnhs_b200 import ops  # noqa: E402
from lgcnhs_b200.trainer import FusedBPRTrainer  # noqa: E402
from model.LightGCN.model import LightGCN  # noqa: E402

dev = torch.device("cuda:0")


def timed(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3


for shape in sys.argv[1:] or ["ml-100k", "ml-1m", "amazon-book"]:
    d = bench.load_shape(shape)
    adj_np, (tr, va, te) = bench.train_adj(d)
    adj = torch.from_numpy(adj_np).to(dev)
    g = torch.Generator().manual_seed(42)
    B, steps = 1024, 60
    pick = torch.randint(len(tr), (steps, B), generator=g)
    users = torch.from_numpy(d.users[tr])[pick].to(dev)
    pos = torch.from_numpy(d.items[tr])[pick].to(dev)
    neg = torch.randint(d.n_items, (steps, B), generator=g).to(dev)
    out = {}
    for dense in ("1", "0"):
        os.environ["LGCNHS_DENSE_BACKWARD"] = dense
        torch.manual_seed(42)
        model = LightGCN(d.n_users, d.n_items, 64, 3).to(dev)
        t = FusedBPRTrainer(model, adj, lr=1e-3, eps_reg=1e-6, deterministic=True)
        for i in range(10):
            t.step(users[i], pos[i], neg[i])
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(10, steps):
            t.step(users[i], pos[i], neg[i])
        e1.record()
        torch.cuda.synchronize()
        out[dense] = (e0.elapsed_time(e1) / (steps - 10) * 1e3, t.X0.clone())
    # the layer alone: X non-zero on one batch's rows
    n = d.n_users + d.n_items
    gr = t.gt
    mask = ops.row_mask_words(n, dev)
    ops.row_mask_batch(mask, users[0], pos[0], neg[0], d.n_users, True)
    X = torch.zeros(n, 64, device=dev)
    rows = torch.cat([users[0], d.n_users + pos[0], d.n_users + neg[0]])
    X[rows] = torch.randn(rows.numel(), 64, device=dev)
    y = torch.empty_like(X)
    t_plain = timed(lambda: gr.spmm(X, X, 1.0, 1.0, out=y))
    y2 = torch.empty_like(X)
    t_mask = timed(lambda: gr.spmm(X, X, 1.0, 1.0, out=y2, src_mask=mask))
    print(f"{shape:12s} step dense {out['1'][0]:7.1f} us  sparse-first-layer {out['0'][0]:7.1f} us  weights identical: "
          f"{torch.equal(out['0'][1], out['1'][1])} | first gradient layer alone: dense {t_plain:6.1f} us  masked {t_mask:6.1f} us  "
          f"equal: {torch.equal(y, y2)}  (nnz {gr.nnz / 1e6:.2f} M, {int(rows.unique().numel())} live rows of {n})", flush=True)
