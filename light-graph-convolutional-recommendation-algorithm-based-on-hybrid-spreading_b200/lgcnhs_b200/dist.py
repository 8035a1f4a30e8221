"""Multi-GPU propagation: one process per GPU (torch.distributed / NCCL for rendezvous and
barriers), rows of A_hat partitioned by non-zeros, per-layer exchange of the new embedding rows.

The reference is single-process; this is the B200-native design of SURVEY.md
§8(e).  Two exchange modes:
  * "p2p"  (default): the all-gather is FUSED INTO THE SpMM — every finished row is stored into all peers'
    replicas through CUDA-IPC-mapped pointers (NVLink stores issued by the kernel that computed the row) — and the
    layer barrier is a one-warp kernel over epoch flags in peer memory (lgc_peer_barrier), so a K-layer call makes
    no NCCL call at all;
  * "p2p-nccl": same stores, one tiny NCCL all-reduce per layer as the barrier (the earlier design, kept for
    comparison);
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


def partition_rows_by_nnz(rowptr: np.ndarray, parts: int, chunk_weight: float = 1.0,
                          long_row: Optional[int] = None) -> np.ndarray:
    """Row boundaries b[0..parts] with ~equal cost per part (rows are never split).  A row costs its non-zeros plus
    one output row; the non-zeros of rows that go through the chunk path (longer than `long_row`, by default the
    threshold a launch of nnz/parts non-zeros would get) count `chunk_weight` times — measured on B200 the chunk path
    runs at ~30 Gnnz/s against ~40 for warp-per-row on launches of this size (tools/spmm_rows_probe.py)."""
    n = rowptr.shape[0] - 1
    deg = np.diff(rowptr.astype(np.int64))
    if long_row is None:
        per_part = max(1.0, float(rowptr[-1] - rowptr[0]) / max(parts, 1))
        long_row = int(min(2048, max(256, 2 ** int(round(np.log2(max(1.0, per_part * 8e-5)))))))
    w = np.where(deg > long_row, chunk_weight, 1.0)
    cost = np.concatenate([[0.0], np.cumsum(deg * w + 1.0)])
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


class _RawCuda:
    """A (rows, cols) fp32 device buffer at a raw address, as a __cuda_array_interface__ object (torch.as_tensor maps it
    zero-copy)."""

    def __init__(self, ptr: int, shape: tuple):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": "<f4", "data": (int(ptr), False), "version": 2,
                                         "strides": None}


class MulticastBuffers:
    """`n_bufs` buffers of `nbytes_each` bytes that exist on EVERY rank at the same offsets and are all bound to ONE
    NVSwitch multicast object: a store to `mc_ptr(i) + off` by any rank is replicated by the switch into every rank's
    copy (NVLS), so the per-layer all-gather of the partitioned SpMM leaves each GPU ONCE instead of once per peer
    (8 x less NVLink egress at 8 GPUs).  Local reads go through the ordinary mapping `local_ptr(i)` of the rank's own
    physical memory.

    Driver API through cuda-python (cuMulticastCreate / AddDevice / BindMem, cuMemCreate / Map / SetAccess); the
    multicast handle travels from rank 0 to the other ranks as a POSIX file descriptor over a Unix-domain socket.
    Raises if the device / driver / fabric does not support multicast — callers fall back to per-peer stores."""

    def __init__(self, nbytes_each: int, n_bufs: int, device: torch.device):
        try:
            from cuda.bindings import driver as cu
        except ImportError:                                   # older cuda-python layout
            from cuda import cuda as cu
        import socket
        import tempfile
        import time

        self.cu = cu
        rank, world = dist.get_rank(), dist.get_world_size()
        dev_index = device.index if device.index is not None else torch.cuda.current_device()

        def ok(res, what):
            err = res[0] if isinstance(res, tuple) else res
            if int(err) != 0:
                raise RuntimeError(f"multicast setup: {what} failed with {err}")
            return res[1] if isinstance(res, tuple) and len(res) == 2 else (res[1:] if isinstance(res, tuple) else None)

        torch.zeros(1, device=device)                         # make sure torch's primary context is current
        cudev = ok(cu.cuDeviceGet(dev_index), "cuDeviceGet")
        sup = ok(cu.cuDeviceGetAttribute(cu.CUdevice_attribute.CU_DEVICE_ATTRIBUTE_MULTICAST_SUPPORTED, cudev), "attribute")
        if not sup:
            raise RuntimeError("multicast setup: CU_DEVICE_ATTRIBUTE_MULTICAST_SUPPORTED is 0")
        fd_type = cu.CUmemAllocationHandleType.CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR
        prop = cu.CUmulticastObjectProp()
        prop.numDevices = world
        prop.handleTypes = fd_type
        prop.flags = 0
        prop.size = int(nbytes_each) * n_bufs
        gran = int(ok(cu.cuMulticastGetGranularity(prop, cu.CUmulticastGranularity_flags.CU_MULTICAST_GRANULARITY_RECOMMENDED),
                      "cuMulticastGetGranularity"))
        self.stride = (int(nbytes_each) + gran - 1) // gran * gran
        size = self.stride * n_bufs
        prop.size = size
        # ---- the multicast object: created by rank 0, imported by the others through a passed file descriptor ----
        name = [os.path.join(tempfile.gettempdir(), f"lgcnhs_mc_{os.getpid()}_{time.time_ns()}.sock")] if rank == 0 else [None]
        dist.broadcast_object_list(name, src=0)
        if rank == 0:
            mc = ok(cu.cuMulticastCreate(prop), "cuMulticastCreate")
            fd = int(ok(cu.cuMemExportToShareableHandle(mc, fd_type, 0), "cuMemExportToShareableHandle"))
            srv = socket.socket(socket.AF_UNIX, socket.SOCK_STREAM)
            srv.bind(name[0])
            srv.listen(world)
            dist.barrier()                                    # the socket exists
            for _ in range(world - 1):
                conn, _a = srv.accept()
                socket.send_fds(conn, [b"mc"], [fd])
                conn.close()
            srv.close()
            os.unlink(name[0])
            os.close(fd)
        else:
            dist.barrier()
            cli = socket.socket(socket.AF_UNIX, socket.SOCK_STREAM)
            cli.connect(name[0])
            _msg, fds, _f, _ad = socket.recv_fds(cli, 16, 1)
            cli.close()
            mc = ok(cu.cuMemImportFromShareableHandle(fds[0], fd_type), "cuMemImportFromShareableHandle")
            os.close(fds[0])
        ok(cu.cuMulticastAddDevice(mc, cudev), "cuMulticastAddDevice")
        dist.barrier()                                        # every device has joined before memory is bound
        # ---- this rank's physical memory, bound into the multicast object at offset 0 ----
        aprop = cu.CUmemAllocationProp()
        aprop.type = cu.CUmemAllocationType.CU_MEM_ALLOCATION_TYPE_PINNED
        aprop.location.type = cu.CUmemLocationType.CU_MEM_LOCATION_TYPE_DEVICE
        aprop.location.id = dev_index
        aprop.requestedHandleTypes = fd_type
        mem = ok(cu.cuMemCreate(size, aprop, 0), "cuMemCreate")
        ok(cu.cuMulticastBindMem(mc, 0, mem, 0, size, 0), "cuMulticastBindMem")
        dist.barrier()
        acc = cu.CUmemAccessDesc()
        acc.location.type = cu.CUmemLocationType.CU_MEM_LOCATION_TYPE_DEVICE
        acc.location.id = dev_index
        acc.flags = cu.CUmemAccess_flags.CU_MEM_ACCESS_FLAGS_PROT_READWRITE
        self.mc_va = int(ok(cu.cuMemAddressReserve(size, gran, 0, 0), "cuMemAddressReserve (mc)"))
        ok(cu.cuMemMap(self.mc_va, size, 0, mc, 0), "cuMemMap (mc)")
        ok(cu.cuMemSetAccess(self.mc_va, size, [acc], 1), "cuMemSetAccess (mc)")
        self.uc_va = int(ok(cu.cuMemAddressReserve(size, gran, 0, 0), "cuMemAddressReserve (local)"))
        ok(cu.cuMemMap(self.uc_va, size, 0, mem, 0), "cuMemMap (local)")
        ok(cu.cuMemSetAccess(self.uc_va, size, [acc], 1), "cuMemSetAccess (local)")
        self.size, self.mc, self.mem, self.n_bufs, self.device = size, mc, mem, n_bufs, device
        torch.cuda.synchronize()
        dist.barrier()

    def mc_ptr(self, i: int) -> int:
        return self.mc_va + i * self.stride

    def local_ptr(self, i: int) -> int:
        return self.uc_va + i * self.stride

    def local_tensor(self, i: int, rows: int, cols: int) -> torch.Tensor:
        return torch.as_tensor(_RawCuda(self.local_ptr(i), (rows, cols)), device=self.device)


class PeerGroup:
    """The ranks of one box as a peer-memory group: device buffers shared through CUDA IPC (every rank gets a pointer to
    every rank's copy) and a stream-ordered device barrier over epoch flags in peer memory (lgc_peer_barrier_dev).
    NCCL / gloo is used for the handle exchange only."""

    def __init__(self, device: torch.device):
        assert dist.is_initialized(), "PeerGroup needs an initialised process group"
        self.rank, self.world, self.dev = dist.get_rank(), dist.get_world_size(), device
        self.flags = torch.zeros(64, dtype=torch.int32, device=device)
        self.epoch_dev = torch.zeros(1, dtype=torch.int32, device=device)
        self.flag_ptrs = self.share(self.flags)
        torch.cuda.synchronize()
        dist.barrier()

    def share(self, t: torch.Tensor) -> list:
        """Pointers to every rank's tensor of this call (same shape on all ranks), index = rank."""
        blobs: list = [None] * self.world
        dist.all_gather_object(blobs, _ipc_export(t))
        return [t.data_ptr() if r == self.rank else _ipc_import(blobs[r]) for r in range(self.world)]

    def shared_matrix(self, rows: int, ld: int):
        """A (rows, ld) fp32 matrix that exists on every rank, with the store targets that reach ALL copies: one NVSwitch
        multicast address when available, else every rank's copy through CUDA IPC.  Returns (local tensor, targets)."""
        mc, good = None, 0
        if os.environ.get("LGCNHS_NO_MULTICAST", "0") != "1":
            try:
                mc = MulticastBuffers(rows * ld * 4, 1, self.dev)
                good = 1
            except Exception as e:        # noqa: BLE001
                self.mcast_error = repr(e)[:300]
        flag = torch.tensor([good], device=self.dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if int(flag.item()) == 1:
            self._keep = getattr(self, "_keep", []) + [mc]
            return mc.local_tensor(0, rows, ld), [mc.mc_ptr(0)]
        t = torch.empty((rows, ld), dtype=torch.float32, device=self.dev)
        return t, self.share(t)

    def barrier(self) -> None:
        farr = (C.c_void_p * self.world)(*[C.c_void_p(p) for p in self.flag_ptrs])
        check(lib().lgc_peer_barrier_dev(self.flags.data_ptr(), farr, self.rank, self.world, self.epoch_dev.data_ptr(),
                                         torch.cuda.current_stream().cuda_stream), "peer barrier")


class RowPartitionedPropagation:
    """K-layer propagation + layer mean with the rows of A_hat split over the ranks."""

    def __init__(self, edge_index: torch.Tensor, n_nodes: int, dim: int, mode: str = "p2p",
                 split: Optional[int] = None):
        """split = n_users partitions the user rows and the item rows SEPARATELY (each rank owns one slice of
        both): user rows gather from the small, popularity-skewed item table and item rows from the large user
        table, so their cost per non-zero differs and a single nnz-balanced cut would leave the item-row ranks
        behind."""
        assert dist.is_initialized(), "RowPartitionedPropagation needs an initialised process group"
        self.rank, self.world = dist.get_rank(), dist.get_world_size()
        self.mode = mode
        self.dev = edge_index.device
        self.n, self.dim = n_nodes, dim
        self.g = NormGraph(edge_index, n_nodes)
        rowptr = self.g.rowptr.cpu().numpy()
        # parts[r] = list of (row_begin, row_end) ranges owned by rank r
        if split is None or split <= 0 or split >= n_nodes:
            b = [int(x) for x in partition_rows_by_nnz(rowptr, self.world, chunk_weight=1.35)]
            self.parts = [[(b[r], b[r + 1])] for r in range(self.world)]
        else:
            bu = [int(x) for x in partition_rows_by_nnz(rowptr[: split + 1], self.world)]
            bi = [int(x) + split for x in partition_rows_by_nnz(rowptr[split:] - rowptr[split], self.world)]
            self.parts = [[(bu[r], bu[r + 1]), (bi[r], bi[r + 1])] for r in range(self.world)]
        self.my_parts = [(a, b, self.g.chunk_range(a, b)) for a, b in self.parts[self.rank] if b > a]
        self.r0, self.r1 = self.parts[self.rank][0]
        # exchange path: "p2p" first tries the NVSwitch multicast mapping (one store per row, replicated by the switch) and
        # falls back to per-peer stores through CUDA IPC if any rank cannot set it up (LGCNHS_NO_MULTICAST=1 forces that)
        self.mcast = None
        if mode == "p2p" and self.world > 1 and os.environ.get("LGCNHS_NO_MULTICAST", "0") != "1":
            try:
                mc = MulticastBuffers(n_nodes * dim * 4, 4, self.dev)
                good = 1
            except Exception as e:        # noqa: BLE001  (any setup failure means: use the peer-store path)
                mc, good = None, 0
                self.mcast_error = repr(e)[:300]
            flag = torch.tensor([good], device=self.dev)
            dist.all_reduce(flag, op=dist.ReduceOp.MIN)
            if int(flag.item()) == 1:
                self.mcast = mc
        # replicated activations: two ping-pong layers + TWO result buffers used alternately by successive calls.
        # The returned tensor aliases a buffer that peers write with remote stores; with a single result buffer a
        # faster rank's NEXT call could overwrite it while this rank's consumers still read it.  With two, the
        # buffer of call n is next written by call n+2, and a peer can only get there after passing a layer
        # barrier of call n+1, which this rank publishes stream-ordered AFTER its consumers of call n's result
        # (requirement: consume the result on the stream the propagation was issued on, or clone it).
        if self.mcast is not None:
            self.bufs = [self.mcast.local_tensor(i, n_nodes, dim) for i in range(4)]
            for b in self.bufs:
                b.zero_()
        else:
            self.bufs = [torch.zeros((n_nodes, dim), dtype=torch.float32, device=self.dev) for _ in range(4)]
        self._calls = 0
        self._flag = torch.zeros(1, dtype=torch.float32, device=self.dev)
        self.peer_ptrs: list[list[int]] = []
        self.peer_flag_ptrs: list[int] = []
        self.epoch = 0
        if mode in ("p2p", "p2p-nccl"):
            # epoch flags live in their own allocation (cudaIpc handles cover whole allocations)
            self.flags = torch.zeros(64, dtype=torch.int32, device=self.dev)
            # barrier epoch, incremented by the barrier kernel itself: the sequence can be captured in a CUDA graph
            self.epoch_dev = torch.zeros(1, dtype=torch.int32, device=self.dev)
            blobs = ([None] * 4 if self.mcast is not None else [_ipc_export(b) for b in self.bufs]) + [_ipc_export(self.flags)]
            gathered: list = [None] * self.world
            dist.all_gather_object(gathered, blobs)
            for bi in range(4):
                if self.mcast is not None:
                    self.peer_ptrs.append([self.mcast.mc_ptr(bi)])     # ONE multicast address reaches every replica
                    continue
                ptrs = []
                for r in range(self.world):
                    ptrs.append(self.bufs[bi].data_ptr() if r == self.rank else _ipc_import(gathered[r][bi]))
                self.peer_ptrs.append(ptrs)
            for r in range(self.world):
                self.peer_flag_ptrs.append(self.flags.data_ptr() if r == self.rank else _ipc_import(gathered[r][4]))
        torch.cuda.synchronize()
        dist.barrier()

    def peer_barrier(self):
        """Stream-ordered device barrier over all ranks (epoch flags in peer memory): returns on this rank's stream
        once every rank's preceding work on ITS stream (row stores into our replicas included) has completed."""
        farr = (C.c_void_p * self.world)(*[C.c_void_p(p) for p in self.peer_flag_ptrs])
        check(lib().lgc_peer_barrier_dev(self.flags.data_ptr(), farr, self.rank, self.world, self.epoch_dev.data_ptr(),
                                         torch.cuda.current_stream().cuda_stream), "peer barrier")

    def _layer_barrier(self):
        # stream-ordered: completes only after every rank's SpMM (and its peer stores) has finished
        dist.all_reduce(self._flag)

    def propagate_mean(self, x0: torch.Tensor, layers: int, result: Optional[int] = None,
                       x0_row_mask: Optional[torch.Tensor] = None) -> torch.Tensor:
        """x0 is replicated on every rank; returns the replicated E = mean_l A_hat^l x0.
        result: which of the two result buffers (0 / 1) receives it; default alternates between calls.
        x0_row_mask: x0 is zero outside the rows whose bit is set (ops.row_mask_words): the first layer skips the gathers
        of the zero rows (p2p mode)."""
        g = self.g
        cur = x0
        res = 2 + ((self._calls & 1) if result is None else int(result))
        self._calls += 1
        for l in range(layers):
            last = l == layers - 1
            oi = res if last else (l & 1)
            alpha = 1.0 / (layers + 1) if last else 1.0
            if self.mode == "p2p":
                # ONE mixed launch over this rank's user rows and item rows (longest first across both)
                g.spmm_rows_bcast(cur, x0, alpha, 1.0, self.peer_ptrs[oi], [(a, b) for a, b, _ in self.my_parts],
                                  src_mask=x0_row_mask if l == 0 else None)
                self.peer_barrier()
            elif self.mode == "p2p-nccl":
                for a, b, ch in self.my_parts:
                    g.spmm_bcast(cur, x0, alpha, 1.0, self.peer_ptrs[oi], a, b, ch)
                self._layer_barrier()
            else:
                out = self.bufs[oi]
                for a, b, ch in self.my_parts:
                    g.spmm(cur, x0, alpha, 1.0, out=out, row_begin=a, row_end=b, chunks=ch)
                for r in range(self.world):
                    for a, b in self.parts[r]:
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
