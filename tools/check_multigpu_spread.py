"""torchrun --nproc-per-node N tools/check_multigpu_spread.py : hybrid spreading sharded with NO
data-path collective — rank r builds the item-column block G[:, J_r] and scores the user-row
block U_r; the blocks must equal the single-GPU result bit for bit, and the gathered top-k ids too."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import numpy as np, torch  # noqa: E402
import torch.distributed as dist  # noqa: E402
from lgcnhs_b200 import ops  # noqa: E402
from lgcnhs_b200.dist import init_dist  # noqa: E402

rank, world, local = init_dist()
dev = torch.device(f"cuda:{local}")
torch.cuda.set_device(dev)
shape = sys.argv[1] if len(sys.argv) > 1 else "ml-1m"
d = bench.load_shape(shape, rank, lambda: dist.barrier())
tr, va, _ = d.split()
sel = np.concatenate([tr, va])
eng = ops.SpreadingEngine(d.n_users, d.n_items, torch.from_numpy(d.users[sel]).to(dev), torch.from_numpy(d.items[sel]).to(dev))
U, M = d.n_users, d.n_items
# single-GPU reference on every rank
G = eng.general_w().clone()
idx_ref, val_ref = eng.recommend(0.3, 20)
# sharded: item block of G, user block of F / top-k
jb = [M * r // world for r in range(world + 1)]
ub = [U * r // world for r in range(world + 1)]
Gblk = eng.general_w(item_range=(jb[rank], jb[rank + 1]))
ok = torch.equal(Gblk, G[:, jb[rank]:jb[rank + 1]])
torch.cuda.synchronize(); dist.barrier(); t0 = time.perf_counter()
for _ in range(10):
    idx, val = eng.recommend(0.3, 20, user_range=(ub[rank], ub[rank + 1]))
torch.cuda.synchronize(); dist.barrier(); dt = (time.perf_counter() - t0) / 10
ok &= torch.equal(idx, idx_ref[ub[rank]:ub[rank + 1]]) and torch.equal(val, val_ref[ub[rank]:ub[rank + 1]])
# the only exchange: gather of the (U, k) ids
parts = [torch.empty((ub[r + 1] - ub[r], 20), dtype=torch.int64, device=dev) for r in range(world)]
dist.all_gather(parts, idx.contiguous()) if len({p.shape for p in parts}) == 1 else [dist.broadcast(parts[r] if r != rank else idx.contiguous(), src=r) for r in range(world)]
print(f"rank {rank}/{world}: G block [{jb[rank]},{jb[rank+1]}) and user block [{ub[rank]},{ub[rank+1]}) bit-identical={ok}; "
      f"sharded lambda-step {dt*1e3:.3f} ms -> {U/dt:.0f} users/s aggregate", flush=True)
# fused GEMM + all-gather: the symmetric schedule dealt over the ranks, every rank ends with the full G, bit-identical
from lgcnhs_b200.dist import PeerGroup  # noqa: E402
group = PeerGroup(dev)
Gall, shared = eng.general_w_allgather(group)
torch.cuda.synchronize()
same_all = torch.equal(Gall, G)
shared[0].fill_(float("nan"))
torch.cuda.synchronize()
dist.barrier()
Gall, shared = eng.general_w_allgather(group, shared=shared)
torch.cuda.synchronize()
same_all &= torch.equal(Gall, G)
ok &= same_all
print(f"rank {rank}/{world}: fused GEMM + all-gather G ({'multicast' if len(shared[1]) == 1 else 'peer stores'}) bit-identical to "
      f"the single-GPU G: {same_all}", flush=True)
flag = torch.tensor([1.0 if ok else 0.0], device=dev)
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
if rank == 0:
    print("MULTIGPU_SPREAD_OK" if flag.item() == 1.0 else "MULTIGPU_SPREAD_MISMATCH", flush=True)
dist.destroy_process_group()
