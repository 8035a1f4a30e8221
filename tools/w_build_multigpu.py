"""torchrun --nproc-per-node N tools/w_build_multigpu.py [shape] : only the W-build leg of bench.py (BASELINE config 5,
fused symmetric GEMM + all-gather) on N GPUs; rank 0 prints its JSON object."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402
from lgcnhs_b200.dist import init_dist  # noqa: E402

rank, world, local = init_dist()
dev = torch.device(f"cuda:{local}")
torch.cuda.set_device(dev)
shape = sys.argv[1] if len(sys.argv) > 1 else "ml-20m"
d = bench.load_shape(shape, rank, lambda: dist.barrier())
out = bench.w_build_leg(d, dev, rank, world, steps=5)
if rank == 0:
    out["n_gpus"] = world
    print(json.dumps(out), flush=True)
dist.barrier()
dist.destroy_process_group()
