"""Multi-GPU propagation: one process per GPU (torch.distributed / NCCL for rendezvous and
barriers), rows of A_hat partitioned by non-zeros, per-layer exchange of the new embedding rows.

The reference is single-process; this is the B200-native design of SURVEY.md
§8(e).  Two exchange modes:
  * "p2p"  (default): the all-gather is FUSED INTO THE SpMM — every finished row is stored into
    all peers' replicas through CUDA-IPC-mapped pointers (NVLink stores issued by the kernel
    that computed the row), followed by one tiny NCCL all-reduce as the layer barrier;
  * "nccl": compute locally, then one broadcast per rank (the library baseline).
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Callable, Optional

import numpy as np
import torch
import torch.distributed as dist

from ._lib import check, lib
from .ops import NormGraph


def init_dist(backend: Optional[str] = None):
    """(rank, world, local_rank); initialises the process group when launched under torchrun."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29511")
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local)
        kw = {"device_id": torch.device("cuda", local)} if backend == "nccl" else {}
        dist.init_process_group(backend=backend, rank=rank, world_size=world, **kw)
    return rank, world, local


def partition_rows_by_nnz(rowptr: np.ndarray, parts: int) -> np.ndarray:
    """Row boundaries b[0..parts] with ~equal non-zeros per part (rows are never split)."""
    n = rowptr.shape[0] - 1
    nnz = int(rowptr[-1])
    # a row costs its non-zeros plus one output row (~1/4 of a non-zero's traffic per 64-wide row)
    cost = rowptr.astype(np.int64) + np.arange(n + 1, dtype=np.int64)
    targets = cost[-1] * np.arange(1, parts, dtype=np.float64) / parts
    cuts = np.searchsorted(cost, targets, side="left")
    b = np.concatenate([[0], cuts, [n]]).astype(np.int64)
    return np.maximum.accumulate(b)


def _ipc_export(t: torch.Tensor) -> bytes:
    buf = (C.c_uint8 * 72)()
    check(lib().lgc_ipc_get_handle(t.data_ptr(), buf), "ipc get handle")
    return bytes(buf)


def _ipc_import(blob: bytes) -> int:
    buf = (C.c_uint8 * 72).from_buffer_copy(blob)
    out = C.c_void_p(0)
    check(lib().lgc_ipc_open_handle(buf, C.byref(out)), "ipc open handle")
    return int(out.value)


class RowPartitionedPropagation:
    """K-layer propagation + layer mean with the rows of A_hat split over the ranks."""

    def __init__(self, edge_index: torch.Tensor, n_nodes: int, dim: int, mode: str = "p2p"):
        assert dist.is_initialized(), "RowPartitionedPropagation needs an initialised process group"
        self.rank, self.world = dist.get_rank(), dist.get_world_size()
        self.mode = mode
        self.dev = edge_index.device
        self.n, self.dim = n_nodes, dim
        self.g = NormGraph(edge_index, n_nodes)
        bounds = partition_rows_by_nnz(self.g.rowptr.cpu().numpy(), self.world)
        self.bounds = [int(b) for b in bounds]
        self.r0, self.r1 = self.bounds[self.rank], self.bounds[self.rank + 1]
        self.chunks = self.g.chunk_range(self.r0, self.r1)
        # replicated activations: two ping-pong layers + the result
        self.bufs = [torch.zeros((n_nodes, dim), dtype=torch.float32, device=self.dev) for _ in range(3)]
        self._flag = torch.zeros(1, dtype=torch.float32, device=self.dev)
        self.peer_ptrs: list[list[int]] = []
        if mode == "p2p":
            blobs = [_ipc_export(b) for b in self.bufs]
            gathered: list = [None] * self.world
            dist.all_gather_object(gathered, blobs)
            for bi in range(3):
                ptrs = []
                for r in range(self.world):
                    ptrs.append(self.bufs[bi].data_ptr() if r == self.rank else _ipc_import(gathered[r][bi]))
                self.peer_ptrs.append(ptrs)
        dist.barrier()

    def _layer_barrier(self):
        # stream-ordered: completes only after every rank's SpMM (and its peer stores) has finished
        dist.all_reduce(self._flag)

    def propagate_mean(self, x0: torch.Tensor, layers: int) -> torch.Tensor:
        """x0 is replicated on every rank; returns the replicated E = mean_l A_hat^l x0."""
        g = self.g
        cur = x0
        for l in range(layers):
            last = l == layers - 1
            oi = 2 if last else (l & 1)
            alpha = 1.0 / (layers + 1) if last else 1.0
            if self.mode == "p2p":
                g.spmm_bcast(cur, x0, alpha, 1.0, self.peer_ptrs[oi], self.r0, self.r1, self.chunks)
                self._layer_barrier()
            else:
                out = self.bufs[oi]
                g.spmm(cur, x0, alpha, 1.0, out=out, row_begin=self.r0, row_end=self.r1, chunks=self.chunks)
                for r in range(self.world):
                    a, b = self.bounds[r], self.bounds[r + 1]
                    if b > a:
                        dist.broadcast(out[a:b], src=r)
            cur = self.bufs[oi]
        return cur if layers > 0 else x0


def emulate_row_partitioned(rowptr: np.ndarray, layer_fn: Callable[[int, int], torch.Tensor], n_rows: int, dim: int):
    """Host-side logic of the nccl mode on any backend (used by the gloo CPU tests): each rank
    computes rows [r0, r1) with `layer_fn`, then the ranks broadcast their row blocks."""
    rank, world = dist.get_rank(), dist.get_world_size()
    bounds = partition_rows_by_nnz(rowptr, world)
    out = torch.zeros((n_rows, dim), dtype=torch.float32)
    r0, r1 = int(bounds[rank]), int(bounds[rank + 1])
    if r1 > r0:
        out[r0:r1] = layer_fn(r0, r1)
    for r in range(world):
        a, b = int(bounds[r]), int(bounds[r + 1])
        if b > a:
            blk = out[a:b].contiguous()
            dist.broadcast(blk, src=r)
            out[a:b] = blk
    return out, bounds
