"""torchrun --nproc-per-node N tools/check_multigpu.py : row-partitioned propagation (fused
SpMM + peer-store all-gather, and the NCCL-broadcast mode) must equal the single-GPU result on every
rank to fp32 round-off (same kernel; a row's summation path depends on the launch size), and bit for bit when the
warp-per-row threshold is pinned (LGCNHS_PIN_LONG_ROW=1)."""
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
if os.environ.get("LGCNHS_PIN_LONG_ROW") == "1":
    from lgcnhs_b200._lib import lib
    lib().lgc_spmm_long_row(256)
ref = ops.NormGraph(adj, n).propagate_mean(x0, 3)
ok = True
for mode, split in (("p2p", d.n_users), ("p2p", None), ("p2p-nccl", d.n_users), ("nccl", d.n_users)):
    prop = RowPartitionedPropagation(adj, n, 64, mode=mode, split=split)
    for it in range(3):
        E = prop.propagate_mean(x0, 3)
    torch.cuda.synchronize()
    same = bool((E - ref).abs().max() <= 2e-6 * ref.abs().max())   # per-launch summation paths: equal to round-off
    ok &= same
    print(f"rank {rank}/{world} mode {mode} (multicast={prop.mcast is not None}{'' if prop.mcast is not None or mode != 'p2p' else ' [' + getattr(prop, 'mcast_error', 'disabled') + ']'}): parts {prop.parts[rank]} equal_to_single_gpu={same} "
          f"max|diff|={(E - ref).abs().max().item():.3e}", flush=True)
    dist.barrier()
flag = torch.tensor([1.0 if ok else 0.0], device=dev)
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
if rank == 0:
    print("MULTIGPU_OK" if flag.item() == 1.0 else "MULTIGPU_MISMATCH", flush=True)
dist.destroy_process_group()
