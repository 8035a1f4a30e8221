"""CUDA-event timing of the fused BPR training step, eager launches vs the captured CUDA graph (GPU box)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench  # noqa: E402
import _stub_const  # noqa: E402
import torch  # noqa: E402

_stub_const.install()
from lgcnhs_b200.trainer import FusedBPRTrainer  # noqa: E402
from model.LightGCN.model import LightGCN  # noqa: E402

dev = torch.device("cuda:0")
for shape in sys.argv[1:] or ["ml-100k", "douban", "ml-1m", "amazon-book"]:
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
    for mode in (False, True):
        torch.manual_seed(42)
        model = LightGCN(d.n_users, d.n_items, 64, 3).to(dev)
        tr_ = FusedBPRTrainer(model, adj, lr=1e-3, eps_reg=1e-6, graph=mode)
        for i in range(10):
            tr_.step(users[i], pos[i], neg[i])
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(10, steps):
            loss = tr_.step(users[i], pos[i], neg[i])
        e1.record()
        torch.cuda.synchronize()
        out[mode] = (e0.elapsed_time(e1) / (steps - 10) * 1e3, float(loss[0]), model.users_emb.weight.detach().clone())
    same = torch.equal(out[False][2], out[True][2])
    print(f"{shape:12s} eager {out[False][0]:8.1f} us/step   graph {out[True][0]:8.1f} us/step   loss {out[False][1]:.6f} / {out[True][1]:.6f}   weights identical: {same}")
