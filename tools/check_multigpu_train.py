"""torchrun --nproc-per-node N tools/check_multigpu_train.py [shape] : the distributed training step (rows of A_hat
partitioned for both propagations, Adam on owned rows with the new rows pushed to every replica) against the single-GPU
FusedBPRTrainer on the same mini-batches: same loss, same weights to Adam-amplified round-off, and bit-identical
weight replicas on all ranks; then the user-block sharded evaluation against the single-GPU ranking."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import numpy as np, torch  # noqa: E402
import torch.distributed as dist  # noqa: E402
from lgcnhs_b200 import ops  # noqa: E402
from lgcnhs_b200.dist import init_dist  # noqa: E402
from lgcnhs_b200.trainer import FusedBPRTrainer, sharded_topk_layer0  # noqa: E402

rank, world, local = init_dist()
dev = torch.device(f"cuda:{local}")
torch.cuda.set_device(dev)
shape = sys.argv[1] if len(sys.argv) > 1 else "ml-100k"
d = bench.load_shape(shape, rank, lambda: dist.barrier())
adj_np, (tr, va, te) = bench.train_adj(d)
adj = torch.from_numpy(adj_np).to(dev)
bench.install_cfg("LightGCN")
from model.LightGCN.model import LightGCN  # noqa: E402

ok = True
for use_graph in (False, True):
    torch.manual_seed(42)
    m1 = LightGCN(d.n_users, d.n_items, 64, 3).to(dev)
    torch.manual_seed(42)
    m2 = LightGCN(d.n_users, d.n_items, 64, 3).to(dev)
    t1 = FusedBPRTrainer(m1, adj, lr=1e-2, eps_reg=1e-4, graph=use_graph)
    t2 = FusedBPRTrainer(m2, adj, lr=1e-2, eps_reg=1e-4, graph=use_graph, distributed=True)
    g = torch.Generator().manual_seed(9)
    steps = 6
    for s in range(steps):
        u = torch.randint(d.n_users, (512,), generator=g).to(dev)
        p = torch.randint(d.n_items, (512,), generator=g).to(dev)
        n = torch.randint(d.n_items, (512,), generator=g).to(dev)
        l1 = t1.step(u, p, n).clone()
        l2 = t2.step(u, p, n).clone()
        torch.cuda.synchronize()
        same_loss = abs(l1[0].item() - l2[0].item()) <= 1e-5 * abs(l1[0].item()) + 1e-7
        err = (t1.X0 - t2.X0).abs().max().item()
        ok &= same_loss and err <= 5e-4 * 1e-2 * (s + 1)
        if rank == 0:
            print(f"graph={use_graph} step {s}: loss {l1[0].item():.6f} / {l2[0].item():.6f}  max|dW| {err:.3e}", flush=True)
    chk = t2.X0.double().sum().reshape(1)
    lo, hi = chk.clone(), chk.clone()
    dist.all_reduce(lo, op=dist.ReduceOp.MIN)
    dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    ok &= bool(lo.item() == hi.item())
    if rank == 0:
        print(f"graph={use_graph}: weight replicas identical on all ranks: {lo.item() == hi.item()}", flush=True)
    e_tr = torch.from_numpy(np.stack([d.users[tr], d.items[tr]]))
    seen = ops.seen_csr(e_tr[0].to(dev), e_tr[1].to(dev), d.n_users, d.n_items)
    full, mine = sharded_topk_layer0(m2, d.n_users, d.n_items, seen, 20, rank, world)
    ref, _ = ops.score_topk(m2.users_emb.weight.detach().contiguous(), m2.items_emb.weight.detach().contiguous(), 20, seen,
                            want_values=False)
    ok &= bool(torch.equal(full, ref))
    if rank == 0:
        print(f"graph={use_graph}: sharded eval == single-GPU eval: {torch.equal(full, ref)}", flush=True)
    del t1, t2
    dist.barrier()
flag = torch.tensor([1.0 if ok else 0.0], device=dev)
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
if rank == 0:
    print("MULTIGPU_TRAIN_OK" if flag.item() == 1.0 else "MULTIGPU_TRAIN_MISMATCH", flush=True)
dist.destroy_process_group()
