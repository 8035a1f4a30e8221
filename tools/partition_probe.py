"""One GPU, emulating the per-rank launches of the row-partitioned propagation: for P = 2, 4, 8 the mixed launch over rank
r's user + item rows (local stores only) under different gathers-in-flight (lgc_spmm_config) and warp-per-row thresholds
(lgc_spmm_long_row).  Tells how much of the multi-GPU inefficiency is the small launch itself."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import numpy as np, torch  # noqa: E402
from lgcnhs_b200 import ops  # noqa: E402
from lgcnhs_b200._lib import lib  # noqa: E402
from lgcnhs_b200.dist import partition_rows_by_nnz  # noqa: E402

dev = torch.device("cuda:0")
shape = sys.argv[1] if len(sys.argv) > 1 else "ml-20m"
d = bench.load_shape(shape)
adj_np, _ = bench.train_adj(d)
n, U = d.n_users + d.n_items, d.n_users
g = ops.NormGraph(torch.from_numpy(adj_np).to(dev), n)
torch.manual_seed(42)
x0 = (torch.randn(n, 64) * 0.1).to(dev)
out = torch.empty_like(x0)
ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
rowptr = g.rowptr.cpu().numpy()


def timed(fn, reps=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ev[0].record()
    for _ in range(reps):
        fn()
    ev[1].record()
    torch.cuda.synchronize()
    return ev[0].elapsed_time(ev[1]) / reps * 1e3


full = timed(lambda: g.spmm(x0, x0, 1.0, 1.0, out=out))
print(f"{shape}: full launch {full:.1f} us ({g.nnz / full / 1e3:.1f} Gnnz/s)")
for P in (2, 4, 8):
    bu = [int(x) for x in partition_rows_by_nnz(rowptr[: U + 1], P)]
    bi = [int(x) + U for x in partition_rows_by_nnz(rowptr[U:] - rowptr[U], P)]
    for r in (0, P - 1):
        ranges = [(bu[r], bu[r + 1]), (bi[r], bi[r + 1])]
        nnz = sum(int(rowptr[b] - rowptr[a]) for a, b in ranges)
        line = f"  P={P} rank {r}: nnz={nnz / 1e6:5.2f} M ideal {full * nnz / g.nnz:6.1f} us |"
        for un in (2, 4, 8):
            lib().lgc_spmm_config(un)
            for lr in (0, 256, 512, 1024):
                lib().lgc_spmm_long_row(lr)
                t = timed(lambda: g.spmm_rows_bcast(x0, x0, 1.0, 1.0, [out.data_ptr()], ranges))
                line += f" un{un}/lr{lr or 'auto'} {t:6.1f}"
            line += " |"
        lib().lgc_spmm_config(0)
        lib().lgc_spmm_long_row(0)
        print(line, flush=True)
