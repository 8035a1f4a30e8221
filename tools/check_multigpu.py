"""torchrun --nproc-per-node N tools/check_multigpu.py : row-partitioned propagation (fused
SpMM + peer-store all-gather, and the NCCL-broadcast mode) must equal the single-GPU result bit
for bit on every rank (same kernel, same per-row summation order)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402
from lgcnhs_b200 import ops  # noqa: E402
from lgcnhs_b200.dist import RowPartitionedPropagation, init_dist  # noqa: E402

rank, world, local = init_dist()
dev = torch.device(f"cuda:{local}")
torch.cuda.set_device(dev)
shape = sys.argv[1] if len(sys.argv) > 1 else "ml-1m"
d = bench.load_shape(shape, rank, lambda: dist.barrier())
adj_np, _ = bench.train_adj(d)
n = d.n_users + d.n_items
adj = torch.from_numpy(adj_np).to(dev)
torch.manual_seed(42)
x0 = (torch.randn(n, 64) * 0.1).to(dev)
ref = ops.NormGraph(adj, n).propagate_mean(x0, 3)
ok = True
for mode in ("p2p", "nccl"):
    prop = RowPartitionedPropagation(adj, n, 64, mode=mode)
    for it in range(3):
        E = prop.propagate_mean(x0, 3)
    torch.cuda.synchronize()
    same = torch.equal(E, ref)
    ok &= same
    print(f"rank {rank}/{world} mode {mode}: rows [{prop.r0},{prop.r1}) equal_to_single_gpu={same} "
          f"max|diff|={(E - ref).abs().max().item():.3e}", flush=True)
    dist.barrier()
flag = torch.tensor([1.0 if ok else 0.0], device=dev)
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
if rank == 0:
    print("MULTIGPU_OK" if flag.item() == 1.0 else "MULTIGPU_MISMATCH", flush=True)
dist.destroy_process_group()
