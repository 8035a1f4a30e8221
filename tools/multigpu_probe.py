"""torchrun --nproc-per-node N tools/multigpu_probe.py [shape] : where does a partitioned propagation layer spend its time?
Per rank: (a) its row ranges computed with local stores only, (b) the same with every row stored into all replicas
(fused all-gather), (c) a full K-layer call (stores + one device barrier per layer).  Device time, every rank printed."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402
from lgcnhs_b200.dist import RowPartitionedPropagation, init_dist  # noqa: E402

rank, world, local = init_dist()
dev = torch.device(f"cuda:{local}")
torch.cuda.set_device(dev)
shape = sys.argv[1] if len(sys.argv) > 1 else "ml-20m"
d = bench.load_shape(shape, rank, lambda: dist.barrier())
adj_np, _ = bench.train_adj(d)
n = d.n_users + d.n_items
adj = torch.from_numpy(adj_np).to(dev)
torch.manual_seed(42)
x0 = (torch.randn(n, 64) * 0.1).to(dev)
for split in (None, d.n_users):
    prop = RowPartitionedPropagation(adj, n, 64, mode="p2p", split=split)
    g = prop.g
    out = torch.empty_like(x0)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]

    def timed(fn, reps=10):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        dist.barrier()
        ev[0].record()
        for _ in range(reps):
            fn()
        ev[1].record()
        torch.cuda.synchronize()
        return ev[0].elapsed_time(ev[1]) / reps * 1e3

    def local_only():
        for a, b, ch in prop.my_parts:
            g.spmm(x0, x0, 1.0, 1.0, out=out, row_begin=a, row_end=b, chunks=ch)

    def with_stores():
        for a, b, ch in prop.my_parts:
            g.spmm_bcast(x0, x0, 1.0, 1.0, prop.peer_ptrs[0], a, b, ch)

    def mixed_stores():
        g.spmm_rows_bcast(x0, x0, 1.0, 1.0, prop.peer_ptrs[0], [(a, b) for a, b, _ in prop.my_parts])

    def with_barrier():
        with_stores()
        prop.peer_barrier()

    dist.barrier()
    t_local = timed(local_only)
    dist.barrier()
    t_store = timed(with_stores)
    dist.barrier()
    t_mixed = timed(mixed_stores)
    t_bar = timed(with_barrier)
    t_full = timed(lambda: prop.propagate_mean(x0, 3), reps=5) / 3
    nnz = sum(int(g.rowptr[b]) - int(g.rowptr[a]) for a, b, _ in prop.my_parts)
    print(f"rank {rank}/{world} multicast={prop.mcast is not None} split={'users/items' if split else 'single'} parts={[(a, b) for a, b, _ in prop.my_parts]} "
          f"nnz={nnz}: local {t_local:.1f} us ({nnz / t_local / 1e3:.1f} Gnnz/s) | +peer stores {t_store:.1f} us | ONE mixed launch "
          f"{t_mixed:.1f} us | "
          f"+barrier {t_bar:.1f} us | full layer {t_full:.1f} us", flush=True)
    dist.barrier()
    del prop
dist.destroy_process_group()
