"""Torch-tensor level wrappers over the C ABI.

torch is used for device memory, streams and (in dist.py) the NCCL process group only;
every numerical kernel below is a hand-written sm_100a kernel in csrc/.  All functions
require CUDA tensors and raise on anything else — there is no CPU path.
"""
from __future__ import annotations

import ctypes as C
import math
import os
from typing import Optional, Sequence

import torch

from ._lib import LgcnhsError, check, lib

LONG_ROW = 256
CHUNK = 1024


def long_row_for(nnz: int) -> int:
    """Warp-per-row threshold of an SpMM launch over `nnz` non-zeros (power of two in [LONG_ROW, 2048])."""
    t = 2 ** int(round(math.log2(max(1.0, nnz * 8e-5))))
    return int(min(2048, max(LONG_ROW, t)))


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _ptr(t: Optional[torch.Tensor]) -> int:
    return 0 if t is None else t.data_ptr()


def _req(t: torch.Tensor, dtype: torch.dtype, name: str) -> torch.Tensor:
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise LgcnhsError(f"{name}: expected a CUDA tensor (the B200 path has no CPU fallback)")
    if t.dtype != dtype:
        raise LgcnhsError(f"{name}: expected dtype {dtype}, got {t.dtype}")
    if not t.is_contiguous():
        raise LgcnhsError(f"{name}: tensor must be contiguous")
    return t


# --------------------------------------------------------------------------------------------
# (P1-P3) normalised graph + propagation
# --------------------------------------------------------------------------------------------
class NormGraph:
    """D^-1/2 A D^-1/2 in target-keyed CSR, built once per graph (gcn_norm semantics).

    edge_index: (2, nnz) int64 CUDA tensor; row 0 = message source, row 1 = target
    (model/LightGCN/model.py:53,62 — `propagate(edge_index, x, norm)`, flow source_to_target).
    """

    def __init__(self, edge_index: torch.Tensor, n_nodes: int):
        ei = _req(edge_index, torch.int64, "edge_index")
        if ei.dim() != 2 or ei.shape[0] != 2:
            raise LgcnhsError("edge_index must have shape (2, nnz)")
        dev = ei.device
        self.device = dev
        self.n_nodes = int(n_nodes)
        self.nnz = int(ei.shape[1])
        L = lib()
        nnz, n = self.nnz, self.n_nodes
        self.rowptr = torch.empty(n + 1, dtype=torch.int32, device=dev)
        self.colidx = torch.empty(max(nnz, 1), dtype=torch.int32, device=dev)
        self.val = torch.empty(max(nnz, 1), dtype=torch.float32, device=dev)
        self.dinv = torch.empty(n, dtype=torch.float32, device=dev)
        max_chunks = int(L.lgc_csr_max_chunks(nnz))
        self.chunk_row = torch.empty(max_chunks, dtype=torch.int32, device=dev)
        self.chunk_start = torch.empty(max_chunks, dtype=torch.int32, device=dev)
        self.row_chunk_base = torch.empty(n + 1, dtype=torch.int32, device=dev)
        ws_bytes = C.c_size_t(0)
        check(L.lgc_csr_build_workspace_bytes(nnz, n, C.byref(ws_bytes)), "csr workspace")
        ws = torch.empty(ws_bytes.value, dtype=torch.uint8, device=dev)
        n_chunks = C.c_int32(0)
        src, dst = ei[0], ei[1]
        check(L.lgc_csr_build(_ptr(src), _ptr(dst), nnz, n, _ptr(self.rowptr), _ptr(self.colidx), _ptr(self.val),
                              _ptr(self.dinv), _ptr(self.chunk_row), _ptr(self.chunk_start),
                              _ptr(self.row_chunk_base), C.byref(n_chunks), _ptr(ws), ws_bytes.value, _stream()),
              "csr build")
        del ws
        self.n_chunks = int(n_chunks.value)
        self.chunk_row = self.chunk_row[: max(self.n_chunks, 1)].clone()
        self.chunk_start = self.chunk_start[: max(self.n_chunks, 1)].clone()
        self._coop = None
        self._lists: dict = {}
        self._scratch: dict[int, tuple[torch.Tensor, torch.Tensor]] = {}
        self._orders: dict[tuple[int, int], torch.Tensor] = {}
        self._long_rows: dict[tuple[int, int], int] = {}
        self.use_row_order = os.environ.get("LGCNHS_NO_ROW_ORDER", "0") != "1"

    def transposed(self) -> "NormGraph":
        """The STRUCTURAL transpose of this operator: entry (r, c, v) -> (c, r, v), values reused, sources ascending
        per row.  This is the graph the backward pass of propagate needs (dX = A_hat^T dY); re-running gcn_norm on the
        transposed edge list would normalise by the OUT-degrees instead and is only equal for a symmetric graph."""
        n, nnz, dev = self.n_nodes, self.nnz, self.device
        t = object.__new__(NormGraph)
        t.device, t.n_nodes, t.nnz = dev, n, nnz
        deg = (self.rowptr[1:] - self.rowptr[:-1]).to(torch.int64)
        rows = torch.repeat_interleave(torch.arange(n, device=dev, dtype=torch.int64), deg)
        cols = self.colidx[:nnz].to(torch.int64)
        # keyed by the old source, old targets ascending inside: own radix sort of the (source, target) keys; a value is
        # ONE fp32 product of the two dinv factors (edge_val_kernel), so it is recomputed instead of permuted — same bits
        keys = sort_u64(cols * n + rows, bits=max(1, (n * n - 1).bit_length()))
        new_rows, new_cols = keys // n, keys % n
        t.colidx = torch.empty(max(nnz, 1), dtype=torch.int32, device=dev)
        t.val = torch.empty(max(nnz, 1), dtype=torch.float32, device=dev)
        t.colidx[:nnz] = new_cols.to(torch.int32)
        t.val[:nnz] = self.dinv[new_rows] * self.dinv[new_cols]
        tdeg = torch.bincount(cols, minlength=n)
        rowptr = torch.zeros(n + 1, dtype=torch.int64, device=dev)
        torch.cumsum(tdeg, 0, out=rowptr[1:])
        t.rowptr = rowptr.to(torch.int32)
        t.dinv = self.dinv
        # long-row chunk list, same rule as lgc_csr_build: rows longer than LONG_ROW are cut into CHUNK-sized chunks
        nch = torch.where(tdeg > LONG_ROW, (tdeg + CHUNK - 1) // CHUNK, torch.zeros_like(tdeg))
        base = torch.zeros(n + 1, dtype=torch.int64, device=dev)
        torch.cumsum(nch, 0, out=base[1:])
        t.row_chunk_base = base.to(torch.int32)
        t.n_chunks = int(base[-1])
        crow = torch.repeat_interleave(torch.arange(n, device=dev, dtype=torch.int64), nch)
        within = torch.arange(t.n_chunks, device=dev, dtype=torch.int64) - base[:-1][crow]
        t.chunk_row = crow.to(torch.int32) if t.n_chunks else torch.zeros(1, dtype=torch.int32, device=dev)
        t.chunk_start = ((rowptr[:-1][crow] + within * CHUNK).to(torch.int32) if t.n_chunks
                         else torch.zeros(1, dtype=torch.int32, device=dev))
        t._scratch, t._orders, t._long_rows, t._coop, t._lists = {}, {}, {}, None, {}
        t.use_row_order = self.use_row_order
        return t

    def row_order(self, row_begin: int = 0, row_end: Optional[int] = None) -> Optional[torch.Tensor]:
        """Rows of [row_begin, row_end) longest first (int32), cached per range; None when disabled."""
        if not self.use_row_order:
            return None
        row_end = self.n_nodes if row_end is None else row_end
        key = (row_begin, row_end)
        o = self._orders.get(key)
        if o is None:
            rp = self.rowptr[row_begin: row_end + 1].to(torch.int64)
            deg = rp[1:] - rp[:-1]
            # longest first, equal lengths by ascending row: one stable sort of (max_deg - deg) << 32 | local row
            # with the library's own radix sort (lgc_sort_u64)
            mx = int(deg.max()) if deg.numel() else 0
            keys = ((mx - deg) << 32) | torch.arange(deg.numel(), device=deg.device, dtype=torch.int64)
            sort_u64(keys, bits=32 + max(1, mx.bit_length()))
            o = ((keys & 0xFFFFFFFF) + row_begin).to(torch.int32)
            self._orders[key] = o
            # warp-per-row threshold sized to the launch: a lone warp streams ~40 non-zeros per microsecond, so rows
            # up to ~8e-5 x nnz(launch) hide inside the launch when issued first (tools/spmm_rows_probe.py and
            # tools/partition_probe.py, B200: 32 M nnz -> 2048, 16 M -> 1024, 4-8 M -> 512, <= 2.4 M -> 256)
            nnz = int(rp[-1] - rp[0])
            self._long_rows[key] = long_row_for(nnz)
        return o

    def long_row(self, row_begin: int = 0, row_end: Optional[int] = None) -> int:
        return self._long_rows.get((row_begin, self.n_nodes if row_end is None else row_end), LONG_ROW)

    def _scr(self, dim: int):
        s = self._scratch.get(dim)
        if s is None:
            partial = torch.empty(max(self.n_chunks, 1) * dim, dtype=torch.float32, device=self.device)
            counters = torch.zeros(self.n_nodes, dtype=torch.int32, device=self.device)
            s = (partial, counters)
            self._scratch[dim] = s
        return s

    def chunk_range(self, row_begin: int, row_end: int) -> tuple[int, int]:
        if row_begin == 0 and row_end == self.n_nodes:
            return 0, self.n_chunks
        b = self.row_chunk_base[[row_begin, row_end]].tolist()
        return int(b[0]), int(b[1])

    def spmm(self, X: torch.Tensor, X0: Optional[torch.Tensor] = None, alpha: float = 1.0, beta: float = 0.0,
             out: Optional[torch.Tensor] = None, row_begin: int = 0, row_end: Optional[int] = None,
             chunks: Optional[tuple[int, int]] = None, src_mask: Optional[torch.Tensor] = None) -> torch.Tensor:
        """out[r] = alpha * (sum_e val[e] X[colidx[e]] + beta X0[r]) for r in [row_begin, row_end).
        src_mask (int32 words, one bit per row of X, see row_mask_words): rows whose bit is clear are all-zero and are not
        gathered — same result, the cost of the (colidx, val) stream only."""
        X = _req(X, torch.float32, "X")
        n, dim = self.n_nodes, int(X.shape[1])
        if X.shape[0] != n:
            raise LgcnhsError(f"X has {X.shape[0]} rows, graph has {n} nodes")
        if X0 is not None:
            _req(X0, torch.float32, "X0")
        if out is None:
            out = torch.empty_like(X)
        _req(out, torch.float32, "out")
        row_end = n if row_end is None else row_end
        cb, ce = chunks if chunks is not None else self.chunk_range(row_begin, row_end)
        partial, counters = self._scr(dim)
        args = (_ptr(self.rowptr), _ptr(self.colidx), _ptr(self.val), _ptr(self.chunk_row),
                _ptr(self.chunk_start), _ptr(self.row_chunk_base), cb, ce, n, dim, row_begin,
                row_end, _ptr(self.row_order(row_begin, row_end)), self.long_row(row_begin, row_end),
                _ptr(X), _ptr(X0), float(alpha), float(beta), _ptr(out), _ptr(partial), _ptr(counters))
        if src_mask is None:
            check(lib().lgc_spmm_layer(*args, _stream()), "spmm layer")
        else:
            check(lib().lgc_spmm_layer_masked(*args, _ptr(_mask_ok(src_mask, n)), _stream()), "spmm layer (masked source)")
        return out

    def spmm_bcast(self, X: torch.Tensor, X0: Optional[torch.Tensor], alpha: float, beta: float,
                   peer_ptrs: Sequence[int], row_begin: int, row_end: int, chunks: tuple[int, int]) -> None:
        """Row range of one layer, every finished row stored into all peers' replicas."""
        X = _req(X, torch.float32, "X")
        n, dim = self.n_nodes, int(X.shape[1])
        partial, counters = self._scr(dim)
        arr = (C.c_void_p * len(peer_ptrs))(*[C.c_void_p(p) for p in peer_ptrs])
        check(lib().lgc_spmm_layer_bcast(_ptr(self.rowptr), _ptr(self.colidx), _ptr(self.val), _ptr(self.chunk_row),
                                         _ptr(self.chunk_start), _ptr(self.row_chunk_base), chunks[0], chunks[1], n,
                                         dim, row_begin, row_end, _ptr(self.row_order(row_begin, row_end)),
                                         self.long_row(row_begin, row_end), _ptr(X), _ptr(X0), float(alpha), float(beta),
                                         arr, len(peer_ptrs), _ptr(partial), _ptr(counters), _stream()),
              "spmm layer bcast")

    COOP_MAX_NNZ = 1_000_000    # graphs up to this many non-zeros run all K layers in one cooperative launch
    COOP_UNIT = 128

    def coop_units(self):
        """Work units of the one-launch K-layer kernel (lgc_propagate_mean_coop), built once per graph: every row is cut
        into pieces of <= COOP_UNIT non-zeros, longest rows first; pieces of a cut row own consecutive partial slots."""
        if self._coop is None:
            dev = self.device
            rp = self.rowptr.to(torch.int64)
            deg = rp[1:] - rp[:-1]
            order = torch.argsort(deg, descending=True, stable=True)
            deg_o = deg[order]
            pieces = torch.clamp((deg_o + self.COOP_UNIT - 1) // self.COOP_UNIT, min=1)
            unit_row = torch.repeat_interleave(order, pieces)
            first_unit = torch.cumsum(pieces, 0) - pieces
            within = torch.arange(int(pieces.sum()), device=dev) - torch.repeat_interleave(first_unit, pieces)
            start = rp[:-1][unit_row] + within * self.COOP_UNIT
            end = torch.minimum(start + self.COOP_UNIT, rp[1:][unit_row])
            is_split = pieces > 1
            n_split = int(is_split.sum())
            split_pieces = pieces[is_split]
            split_first = torch.cumsum(split_pieces, 0) - split_pieces
            slot_of_row = torch.full((self.n_nodes,), -1, dtype=torch.int64, device=dev)
            slot_of_row[order[is_split]] = split_first
            slot = torch.where(slot_of_row[unit_row] >= 0, slot_of_row[unit_row] + within, torch.full_like(within, -1))
            i32 = lambda t: t.to(torch.int32).contiguous()  # noqa: E731
            self._coop = {"unit_row": i32(unit_row), "start": i32(start), "end": i32(end), "slot": i32(slot),
                          "split_row": i32(order[is_split]) if n_split else None,
                          "split_first": i32(split_first) if n_split else None,
                          "split_count": i32(split_pieces) if n_split else None,
                          "n_units": int(unit_row.numel()), "n_split": n_split, "n_partials": int(split_pieces.sum()) if n_split else 0,
                          "partial": {}, "barrier": torch.zeros(2, dtype=torch.int32, device=dev)}
        return self._coop

    def row_list(self, ranges: Sequence[tuple[int, int]]):
        """(rows of the given ranges longest first as one int32 list, warp-per-row threshold of a launch over them, up to
        two chunk ranges) for lgc_spmm_rows_bcast; cached per set of ranges."""
        key = tuple((int(a), int(b)) for a, b in ranges if b > a)
        hit = self._lists.get(key)
        if hit is None:
            rp = self.rowptr.to(torch.int64)
            rows = torch.cat([torch.arange(a, b, device=self.device, dtype=torch.int64) for a, b in key])
            deg = rp[rows + 1] - rp[rows]
            mx = int(deg.max()) if deg.numel() else 0
            keys = ((mx - deg) << 32) | rows
            sort_u64(keys, bits=32 + max(1, mx.bit_length()))
            nnz = int(deg.sum())
            chunks = [self.chunk_range(a, b) for a, b in key][:2]
            while len(chunks) < 2:
                chunks.append((0, 0))
            if len(key) > 2:
                raise LgcnhsError("row_list: at most two row ranges per launch")
            hit = ((keys & 0xFFFFFFFF).to(torch.int32), long_row_for(nnz), chunks)
            self._lists[key] = hit
        return hit

    def spmm_rows_bcast(self, X: torch.Tensor, X0: Optional[torch.Tensor], alpha: float, beta: float,
                        peer_ptrs: Sequence[int], ranges: Sequence[tuple[int, int]],
                        src_mask: Optional[torch.Tensor] = None) -> None:
        """One mixed launch over up to two row ranges, every finished row stored into all replicas.  src_mask: as in spmm."""
        X = _req(X, torch.float32, "X")
        n, dim = self.n_nodes, int(X.shape[1])
        rows, long_row, ch = self.row_list(ranges)
        partial, counters = self._scr(dim)
        arr = (C.c_void_p * len(peer_ptrs))(*[C.c_void_p(p) for p in peer_ptrs])
        args = (_ptr(self.rowptr), _ptr(self.colidx), _ptr(self.val), _ptr(self.chunk_row),
                _ptr(self.chunk_start), _ptr(self.row_chunk_base), ch[0][0], ch[0][1], ch[1][0],
                ch[1][1], n, dim, _ptr(rows), int(rows.numel()), long_row, _ptr(X), _ptr(X0),
                float(alpha), float(beta), arr, len(peer_ptrs), _ptr(partial), _ptr(counters))
        if src_mask is None:
            check(lib().lgc_spmm_rows_bcast(*args, _stream()), "spmm rows bcast")
        else:
            check(lib().lgc_spmm_rows_bcast_masked(*args, _ptr(_mask_ok(src_mask, n)), _stream()),
                  "spmm rows bcast (masked source)")

    def propagate_mean(self, X0: torch.Tensor, n_layers: int, out: Optional[torch.Tensor] = None,
                       tmp: Optional[tuple[torch.Tensor, torch.Tensor]] = None, coop: Optional[bool] = None,
                       x0_row_mask: Optional[torch.Tensor] = None) -> torch.Tensor:
        """E = mean_l(A_hat^l X0), l = 0..n_layers (model/LightGCN/model.py:56-69).
        x0_row_mask (see row_mask_words / row_mask_batch): X0 is zero outside the rows whose bit is set — the first layer
        (the only one that reads X0 as its source) then skips the gathers of the zero rows; used for dL/dE in training.
        coop=True (or LGCNHS_COOP=1 for graphs with <= COOP_MAX_NNZ non-zeros): all layers in ONE cooperative launch with
        grid barriers between them.  Opt-in: measured on B200 (tools/coop_probe.py, profiles/r2_coop_probe.txt) the grid
        barriers cost more than the launches they replace at every configured shape."""
        X0 = _req(X0, torch.float32, "X0")
        n, dim = self.n_nodes, int(X0.shape[1])
        if X0.shape[0] != n:
            raise LgcnhsError(f"X0 has {X0.shape[0]} rows, graph has {n} nodes")
        if out is None:
            out = torch.empty_like(X0)
        if tmp is None:
            tmp = (torch.empty_like(X0), torch.empty_like(X0))
        if coop is None:
            coop = 0 < self.nnz <= self.COOP_MAX_NNZ and n_layers >= 1 and os.environ.get("LGCNHS_COOP", "0") == "1"
        if x0_row_mask is not None and n_layers >= 1:
            partial, counters = self._scr(dim)
            check(lib().lgc_propagate_mean_masked(_ptr(self.rowptr), _ptr(self.colidx), _ptr(self.val), _ptr(self.chunk_row),
                                                  _ptr(self.chunk_start), _ptr(self.row_chunk_base), self.n_chunks, n, dim,
                                                  int(n_layers), _ptr(self.row_order()), self.long_row(), _ptr(X0), _ptr(out),
                                                  _ptr(tmp[0]), _ptr(tmp[1]), _ptr(partial), _ptr(counters),
                                                  _ptr(_mask_ok(x0_row_mask, n)), _stream()), "propagate_mean (masked source)")
            return out
        if coop:
            cu = self.coop_units()
            part = cu["partial"].get(dim)
            if part is None:
                part = cu["partial"][dim] = torch.empty(max(cu["n_partials"], 1) * dim, dtype=torch.float32, device=self.device)
            check(lib().lgc_propagate_mean_coop(_ptr(self.rowptr), _ptr(self.colidx), _ptr(self.val), _ptr(cu["unit_row"]),
                                                _ptr(cu["start"]), _ptr(cu["end"]), _ptr(cu["slot"]), cu["n_units"],
                                                _ptr(cu["split_row"]), _ptr(cu["split_first"]), _ptr(cu["split_count"]),
                                                cu["n_split"], n, dim, int(n_layers), _ptr(X0), _ptr(out), _ptr(tmp[0]),
                                                _ptr(tmp[1]), _ptr(part), _ptr(cu["barrier"]), _stream()),
                  "propagate_mean (cooperative)")
            return out
        partial, counters = self._scr(dim)
        check(lib().lgc_propagate_mean(_ptr(self.rowptr), _ptr(self.colidx), _ptr(self.val), _ptr(self.chunk_row),
                                       _ptr(self.chunk_start), _ptr(self.row_chunk_base), self.n_chunks, n, dim,
                                       int(n_layers), _ptr(self.row_order()), self.long_row(), _ptr(X0), _ptr(out),
                                       _ptr(tmp[0]), _ptr(tmp[1]), _ptr(partial), _ptr(counters), _stream()),
              "propagate_mean")
        return out

    # algorithmic bytes of one layer, SURVEY.md §8(d): no-reuse model
    def layer_bytes(self, dim: int) -> int:
        return self.nnz * (4 + 4 + 4 * dim) + self.n_nodes * (4 * dim + 4) + 4

    def layer_bytes_compulsory(self, dim: int) -> int:
        return self.nnz * 8 + (self.n_nodes + 1) * 4 + 2 * self.n_nodes * 4 * dim


def row_mask_words(n_rows: int, device) -> torch.Tensor:
    """An all-clear row mask for n_rows rows: int32 words, bit (r & 31) of word r >> 5 belongs to row r."""
    return torch.zeros(((n_rows + 31) // 32 + 1,), dtype=torch.int32, device=device)


def _mask_ok(mask: torch.Tensor, n_rows: int) -> torch.Tensor:
    if mask.dtype != torch.int32 or not mask.is_cuda or not mask.is_contiguous() or mask.numel() * 32 < n_rows:
        raise LgcnhsError("row mask: contiguous CUDA int32 words covering every row expected (row_mask_words)")
    return mask


def row_mask_batch(mask: torch.Tensor, users: torch.Tensor, pos: torch.Tensor, neg: torch.Tensor, n_users: int,
                   set_bits: bool) -> None:
    """Set (or clear again) the bits of rows users[b], n_users + pos[b], n_users + neg[b]: the non-zero rows of dL/dE."""
    check(lib().lgc_row_mask_batch(_ptr(mask), _ptr(_req(users, torch.int64, "users")), _ptr(_req(pos, torch.int64, "pos")),
                                   _ptr(_req(neg, torch.int64, "neg")), int(users.numel()), int(n_users), 1 if set_bits else 0,
                                   _stream()), "row mask batch")


def probe_gather_gbs(n_rows: int, dim: int = 64, n_gathers: int = 32_000_000, reps: int = 5, device=None) -> dict:
    """Measured throughput of random whole-row gathers from an (n_rows, dim) fp32 table (lgc_probe_gather): the L2
    gather peak of this GPU for this table size.  Returns {"gbs", "us", "rows", "table_mb"}; best of `reps`."""
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else device
    table = torch.randn((n_rows, dim), dtype=torch.float32, device=dev)
    out = torch.empty(int(lib().lgc_probe_gather_threads()), dtype=torch.float32, device=dev)
    done = C.c_int64(0)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    best = float("inf")
    for r in range(reps + 1):
        ev[0].record()
        check(lib().lgc_probe_gather(_ptr(table), n_rows, dim, n_gathers, 1234 + r, _ptr(out), C.byref(done), _stream()),
              "probe gather")
        ev[1].record()
        torch.cuda.synchronize()
        if r > 0:
            best = min(best, ev[0].elapsed_time(ev[1]))
    nbytes = done.value * dim * 4
    return {"gbs": nbytes / (best * 1e-3) / 1e9, "us": best * 1e3, "rows": int(done.value),
            "table_mb": n_rows * dim * 4 / 1e6}


# --------------------------------------------------------------------------------------------
# (P4/P6/P7) BPR + Adam
# --------------------------------------------------------------------------------------------
_bpr_scratch: dict[torch.device, torch.Tensor] = {}


def _bpr_scr(dev: torch.device) -> torch.Tensor:
    s = _bpr_scratch.get(dev)
    if s is None:
        s = torch.zeros(int(lib().lgc_bpr_scratch_floats(0)), dtype=torch.float32, device=dev)
        _bpr_scratch[dev] = s
    return s


def bpr_fwd_bwd(E: torch.Tensor, X0: torch.Tensor, n_users: int, n_items: int, users: torch.Tensor,
                pos: torch.Tensor, neg: torch.Tensor, eps: float, gE: Optional[torch.Tensor] = None,
                gX0: Optional[torch.Tensor] = None, grad_scale: float = 1.0,
                loss_out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Returns loss_out = [total, bpr]; if gE/gX0 are given the row gradients are scatter-added."""
    E = _req(E, torch.float32, "E")
    X0 = _req(X0, torch.float32, "X0")
    dim = int(E.shape[1])
    for t, nm in ((users, "users"), (pos, "pos"), (neg, "neg")):
        _req(t, torch.int64, nm)
    loss = torch.empty(2, dtype=torch.float32, device=E.device) if loss_out is None else loss_out
    check(lib().lgc_bpr_fwd_bwd(_ptr(E), _ptr(X0), n_users, n_items, dim, _ptr(users), _ptr(pos), _ptr(neg),
                                int(users.numel()), float(eps), float(grad_scale), _ptr(loss), _ptr(gE), _ptr(gX0),
                                _ptr(_bpr_scr(E.device)), _stream()), "bpr")
    return loss


_bpr_det_ws: dict = {}


def bpr_fwd_bwd_det(E: torch.Tensor, X0: torch.Tensor, n_users: int, n_items: int, users: torch.Tensor, pos: torch.Tensor,
                    neg: torch.Tensor, eps: float, gE: torch.Tensor, gX0: torch.Tensor, grad_scale: float = 1.0,
                    loss_out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """bpr_fwd_bwd with the deterministic (atomic-free) gradient scatter: bit-identical gradients from run to run."""
    E = _req(E, torch.float32, "E")
    X0 = _req(X0, torch.float32, "X0")
    dim, batch = int(E.shape[1]), int(users.numel())
    for t, nm in ((users, "users"), (pos, "pos"), (neg, "neg")):
        _req(t, torch.int64, nm)
    loss = torch.empty(2, dtype=torch.float32, device=E.device) if loss_out is None else loss_out
    nbytes = int(lib().lgc_bpr_det_workspace_bytes(batch, dim))
    ws = _bpr_det_ws.get((E.device, batch, dim))
    if ws is None:
        ws = _bpr_det_ws[(E.device, batch, dim)] = torch.empty(nbytes, dtype=torch.uint8, device=E.device)
    check(lib().lgc_bpr_fwd_bwd_det(_ptr(E), _ptr(X0), n_users, n_items, dim, _ptr(users), _ptr(pos), _ptr(neg), batch,
                                    float(eps), float(grad_scale), _ptr(loss), _ptr(gE), _ptr(gX0), _ptr(_bpr_scr(E.device)),
                                    _ptr(ws), nbytes, _stream()), "bpr (deterministic)")
    return loss


def bpr_rows(rows: Sequence[torch.Tensor], eps: float, grads: Optional[Sequence[torch.Tensor]] = None,
             grad_scale: float = 1.0) -> torch.Tensor:
    """rows = (u_f, u_0, p_f, p_0, n_f, n_0), each (B, dim) fp32."""
    rows = [_req(t, torch.float32, "bpr row block") for t in rows]
    B, dim = int(rows[0].shape[0]), int(rows[0].shape[1])
    loss = torch.empty(2, dtype=torch.float32, device=rows[0].device)
    g = [None] * 6 if grads is None else list(grads)
    check(lib().lgc_bpr_rows(*[_ptr(t) for t in rows], B, dim, float(eps), float(grad_scale), _ptr(loss),
                             *[_ptr(t) for t in g], _ptr(_bpr_scr(rows[0].device)), _stream()), "bpr rows")
    return loss


def adam_step(param: torch.Tensor, grad: torch.Tensor, exp_avg: torch.Tensor, exp_avg_sq: torch.Tensor, step: int,
              lr: float, beta1: float = 0.9, beta2: float = 0.999, eps: float = 1e-8) -> None:
    for t, nm in ((param, "param"), (grad, "grad"), (exp_avg, "exp_avg"), (exp_avg_sq, "exp_avg_sq")):
        _req(t, torch.float32, nm)
    bc1 = 1.0 - beta1 ** step
    bc2_sqrt = math.sqrt(1.0 - beta2 ** step)
    check(lib().lgc_adam_step(_ptr(param), _ptr(grad), _ptr(exp_avg), _ptr(exp_avg_sq), int(param.numel()), float(lr),
                              float(beta1), float(beta2), float(eps), float(bc1), float(bc2_sqrt), _stream()), "adam")


def adam_hyper_step(step_dev: torch.Tensor, lr_dev: torch.Tensor, beta1: float, beta2: float, hyper_dev: torch.Tensor) -> None:
    """++step (device int64) and hyper = {lr / (1 - beta1^step), sqrt(1 - beta2^step)} on the device."""
    check(lib().lgc_adam_hyper_step(_ptr(step_dev), _ptr(lr_dev), float(beta1), float(beta2), _ptr(hyper_dev), _stream()),
          "adam hyper")


def adam_step_dev(param: torch.Tensor, grad: torch.Tensor, exp_avg: torch.Tensor, exp_avg_sq: torch.Tensor, beta1: float,
                  beta2: float, eps: float, hyper_dev: torch.Tensor) -> None:
    for t, nm in ((param, "param"), (grad, "grad"), (exp_avg, "exp_avg"), (exp_avg_sq, "exp_avg_sq")):
        _req(t, torch.float32, nm)
    check(lib().lgc_adam_step_dev(_ptr(param), _ptr(grad), _ptr(exp_avg), _ptr(exp_avg_sq), int(param.numel()), float(beta1),
                                  float(beta2), float(eps), _ptr(hyper_dev), _stream()), "adam (device hyper)")


# --------------------------------------------------------------------------------------------
# (P8/S4) scores and top-k
# --------------------------------------------------------------------------------------------
def score_block(Xu: torch.Tensor, Xi: torch.Tensor, u0: int, u1: int, seen: Optional[tuple] = None,
                fill: float = -1024.0, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    Xu = _req(Xu, torch.float32, "Xu")
    Xi = _req(Xi, torch.float32, "Xi")
    M, dim = int(Xi.shape[0]), int(Xi.shape[1])
    if out is None:
        out = torch.empty((u1 - u0, M), dtype=torch.float32, device=Xu.device)
    sp, si = (None, None) if seen is None else seen
    check(lib().lgc_score_block(_ptr(Xu), _ptr(Xi), u0, u1, M, dim, _ptr(sp), _ptr(si), float(fill), _ptr(out),
                                int(out.stride(0)), _stream()), "score block")
    return out


_tc_ws: dict = {}


def score_topk(Xu: torch.Tensor, Xi: torch.Tensor, k: int, seen: Optional[tuple] = None, fill: float = -1024.0,
               exclude_seen: bool = False, mul: Optional[torch.Tensor] = None, u0: int = 0, u1: Optional[int] = None,
               want_values: bool = True, precise: Optional[bool] = None):
    """Fused  top-k_i ( (seen ? fill : <Xu[u], Xi[i]>) * mul[u, i] )  for users [u0, u1): the score matrix is never
    materialised (reference: matmul + index_put(-1024) + topk, model/LightGCN/recommend.py:86-114).

    Two kernels behind one contract:
      * tensor cores (default when dim in {32, 64}, k <= 32 and more than 128 users): tcgen05 kind::tf32 with a 3xTF32
        split of both operands (lgc_score_topk_tc) — scores within ~1e-6 relative of an fp32 SGEMM, ids identical
        except at float near-ties;
      * precise=True (or LGCNHS_SCORE_FP32=1, or any shape the tensor-core kernel does not take): packed fp32 FMA
        (lgc_score_topk), bit-identical to the scalar fp32 dot product."""
    Xu = _req(Xu, torch.float32, "Xu")
    Xi = _req(Xi, torch.float32, "Xi")
    M, dim = int(Xi.shape[0]), int(Xi.shape[1])
    u1 = int(Xu.shape[0]) if u1 is None else u1
    sp, si = (None, None) if seen is None else seen
    if mul is not None:
        if mul.dtype != torch.float32 or not mul.is_cuda or mul.stride(1) != 1 or mul.shape[0] != u1 - u0 or mul.shape[1] != M:
            raise LgcnhsError("score_topk: mul must be a CUDA fp32 (u1-u0, n_items) matrix with unit column stride")
    idx = torch.empty((u1 - u0, k), dtype=torch.int64, device=Xu.device)
    val = torch.empty((u1 - u0, k), dtype=torch.float32, device=Xu.device) if want_values else None
    if precise is None:
        precise = os.environ.get("LGCNHS_SCORE_FP32", "0") == "1"
    if not precise and dim in (32, 64) and k <= 32 and (u1 - u0) > 128:
        n_total = int(Xu.shape[0])
        nbytes = int(lib().lgc_score_topk_tc_workspace_bytes(n_total, M, u1 - u0, dim))
        ws = _tc_ws.get(Xu.device)
        if ws is None or ws.numel() < nbytes:
            ws = _tc_ws[Xu.device] = torch.empty(nbytes, dtype=torch.uint8, device=Xu.device)
        check(lib().lgc_score_topk_tc(_ptr(Xu), _ptr(Xi), n_total, u0, u1, M, dim, _ptr(sp), _ptr(si), float(fill),
                                      int(exclude_seen), _ptr(mul), int(mul.stride(0)) if mul is not None else 0, int(k),
                                      _ptr(idx), _ptr(val), _ptr(ws), nbytes, _stream()), "score_topk_tc")
        return idx, val
    check(lib().lgc_score_topk(_ptr(Xu), _ptr(Xi), u0, u1, M, dim, _ptr(sp), _ptr(si), float(fill), int(exclude_seen),
                               _ptr(mul), int(mul.stride(0)) if mul is not None else 0, int(k), _ptr(idx), _ptr(val),
                               _stream()), "score_topk")
    return idx, val


class ExclusionMask:
    """Bit-packed (n_rows x n_cols) matrix of entries that must never be recommended."""

    def __init__(self, bits: torch.Tensor, n_rows: int, n_cols: int, stride_bits: int):
        self.bits, self.n_rows, self.n_cols, self.stride_bits = bits, n_rows, n_cols, stride_bits

    @staticmethod
    def from_csr(csr: tuple, n_rows: int, n_cols: int) -> "ExclusionMask":
        ptr, idx = csr
        bits = torch.zeros((n_rows * n_cols + 31) // 32, dtype=torch.int32, device=ptr.device)
        check(lib().lgc_mask_from_csr(_ptr(ptr), _ptr(idx), n_rows, n_cols, n_cols, _ptr(bits), _stream()), "mask from csr")
        return ExclusionMask(bits, n_rows, n_cols, n_cols)

    @staticmethod
    def from_pairs(users: torch.Tensor, items: torch.Tensor, n_rows: int, n_cols: int) -> "ExclusionMask":
        return ExclusionMask.from_csr(seen_csr(users, items, n_rows, n_cols), n_rows, n_cols)


def topk_rows(S: torch.Tensor, k: int, excl: Optional[ExclusionMask] = None, row_offset: int = 0,
              want_values: bool = True):
    """Row-wise top-k (sorted, ties -> larger index).  Row r of S is row `row_offset + r` of the mask."""
    if S.dtype != torch.float32 or not S.is_cuda or S.dim() != 2 or S.stride(1) != 1:
        raise LgcnhsError("topk: S must be a CUDA fp32 matrix with unit column stride")
    rows, cols = int(S.shape[0]), int(S.shape[1])
    if excl is not None and (excl.n_cols != cols or row_offset + rows > excl.n_rows):
        raise LgcnhsError("topk: exclusion mask does not cover the score block")
    idx = torch.empty((rows, k), dtype=torch.int64, device=S.device)
    val = torch.empty((rows, k), dtype=torch.float32, device=S.device) if want_values else None
    check(lib().lgc_topk_rows(_ptr(S), rows, cols, int(S.stride(0)), _ptr(excl.bits) if excl else 0,
                              excl.stride_bits if excl else 0, int(row_offset), int(k), _ptr(idx), _ptr(val), _stream()),
          "topk rows")
    return idx, val


def topk_metrics(rec: torch.Tensor, n_items: int, pos: Optional[tuple] = None, cooc: Optional[torch.Tensor] = None,
                 item_deg: Optional[torch.Tensor] = None, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Raw sums of the six metrics of the (U, k) lists `rec` as a float64[6] DEVICE tensor (see lgc_metrics_topk);
    `metrics_from_sums` turns them into the reference's rounded numbers."""
    rec = _req(rec, torch.int64, "rec")
    U, k = int(rec.shape[0]), int(rec.shape[1])
    if out is None:
        out = torch.empty(6, dtype=torch.float64, device=rec.device)
    pp, pi = (None, None) if pos is None else pos
    if cooc is not None and (cooc.dtype != torch.float32 or cooc.stride(1) != 1 or cooc.shape[0] != n_items):
        raise LgcnhsError("topk_metrics: cooc must be the fp32 (n_items, n_items) matrix A^T A")
    scratch = torch.empty(int(lib().lgc_metrics_scratch_bytes(n_items)), dtype=torch.uint8, device=rec.device)
    check(lib().lgc_metrics_topk(_ptr(rec), U, k, int(n_items), _ptr(pp), _ptr(pi), _ptr(cooc),
                                 int(cooc.stride(0)) if cooc is not None else 0, _ptr(item_deg), _ptr(out), _ptr(scratch),
                                 _stream()), "metrics_topk")
    return out


def metrics_from_sums(sums, n_users: int, k: int) -> dict:
    """float64[6] sums (host) -> the reference's six numbers, rounded to 5 decimals like metrics/*.py."""
    s = [float(x) for x in sums]
    n = max(s[3], 1.0)
    precision, recall, ndcg = round(s[0] / n / k, 5), round(s[1] / n, 5), round(s[2] / n, 5)
    f1 = round(2 * precision * recall / (precision + recall), 5) if precision + recall > 0 else float("nan")
    H = round(1.0 - s[4] / (n_users * (n_users - 1) * k), 5) if n_users > 1 else float("nan")
    I = round(s[5] / (n_users * k * (k - 1)), 5) if k > 1 else float("nan")
    return {"precision": precision, "recall": recall, "f1": f1, "ndcg": ndcg, "H": H, "I": I}


def seen_csr(users: torch.Tensor, items: torch.Tensor, n_users: int, n_items: int):
    """(user, item) pairs -> deduplicated int32 CSR (rowptr, item ids ascending) on device: own radix sort + scan +
    compaction kernels (lgc_seen_csr), standing in for the reference's Python dict / list loops over the pairs
    (utils/trans.py:51-80, model/LightGCN/recommend.py:92-111)."""
    if not users.is_cuda:
        raise LgcnhsError("seen_csr: expected CUDA tensors (the B200 path has no CPU fallback)")
    users = users.to(torch.int64).contiguous()
    items = items.to(torch.int64).contiguous()
    n = int(users.numel())
    dev = users.device
    rowptr = torch.empty(n_users + 1, dtype=torch.int32, device=dev)
    idx = torch.empty(max(n, 1), dtype=torch.int32, device=dev)
    nb = C.c_size_t(0)
    check(lib().lgc_seen_csr_workspace_bytes(n, C.byref(nb)), "seen csr workspace")
    ws = torch.empty(nb.value, dtype=torch.uint8, device=dev)
    n_unique = C.c_int64(0)
    check(lib().lgc_seen_csr(_ptr(users), _ptr(items), n, int(n_users), int(n_items), _ptr(rowptr), _ptr(idx),
                             C.byref(n_unique), _ptr(ws), nb.value, _stream()), "seen csr")
    return rowptr, idx[: n_unique.value].clone() if n_unique.value < n else idx[:n]


def unique_u64(keys: torch.Tensor, bits: int = 64) -> torch.Tensor:
    """Distinct values of non-negative int64 keys, ascending (own radix sort + scan + compaction, lgc_unique_u64); the
    input tensor is clobbered."""
    keys = _req(keys, torch.int64, "keys")
    n = int(keys.numel())
    if n == 0:
        return keys
    out = torch.empty_like(keys)
    nb = C.c_size_t(0)
    check(lib().lgc_unique_u64_workspace_bytes(n, C.byref(nb)), "unique workspace")
    ws = torch.empty(nb.value, dtype=torch.uint8, device=keys.device)
    cnt = C.c_int64(0)
    check(lib().lgc_unique_u64(_ptr(keys), _ptr(out), n, int(bits), C.byref(cnt), _ptr(ws), nb.value, _stream()), "unique")
    return out[: cnt.value]


def sort_u64(keys: torch.Tensor, bits: int = 64) -> torch.Tensor:
    """Stable ascending radix sort of non-negative int64 keys on the device (lgc_sort_u64), in place; returns keys."""
    keys = _req(keys, torch.int64, "keys")
    n = int(keys.numel())
    if n == 0:
        return keys
    tmp = torch.empty_like(keys)
    nb = C.c_size_t(0)
    check(lib().lgc_sort_u64_workspace_bytes(n, C.byref(nb)), "sort workspace")
    ws = torch.empty(nb.value, dtype=torch.uint8, device=keys.device)
    check(lib().lgc_sort_u64(_ptr(keys), _ptr(tmp), n, int(bits), _ptr(ws), nb.value, _stream()), "sort")
    return keys


# --------------------------------------------------------------------------------------------
# (S1-S3) spreading GEMMs
# --------------------------------------------------------------------------------------------
def gemm_planes(kind: int, A: torch.Tensor, B: torch.Tensor, M: int, N: int, K: int,
                out: Optional[torch.Tensor] = None, rs: Optional[torch.Tensor] = None,
                cs: Optional[torch.Tensor] = None, scale: float = 1.0, simt: bool = False) -> torch.Tensor:
    """C = rs*cs*scale * sum_p w_p A B_p^T.  A: (M, lda); B: (planes, N, ldb); K-major."""
    dt = torch.bfloat16 if kind == 0 else torch.uint8
    for t, nm in ((A, "A"), (B, "B planes")):
        if not t.is_cuda or t.dtype != dt or t.stride(-1) != 1:
            raise LgcnhsError(f"{nm}: expected a CUDA {dt} tensor with unit inner stride")
    if B.dim() != 3 or A.dim() != 2:
        raise LgcnhsError("A must be (M, lda) and B (planes, N, ldb)")
    planes = int(B.shape[0])
    if out is None:
        ldc = (N + 3) // 4 * 4
        out = torch.empty((M, ldc), dtype=torch.float32, device=A.device)[:, :N]
    fn = lib().hs_gemm_planes_simt if simt else lib().hs_gemm_planes
    check(fn(kind, _ptr(A), int(A.stride(0)), _ptr(B), int(B.stride(1)), int(B.stride(0)), planes, M, N, K,
             _ptr(out), int(out.stride(0)), _ptr(rs), _ptr(cs), float(scale), _stream()), "gemm planes")
    return out


def _pad(x: int, m: int) -> int:
    return (x + m - 1) // m * m


class SpreadingEngine:
    """Device-resident hybrid HeatS/ProbS spreading for one interaction matrix A (train ⊕ val).

    Mirrors the call sequence of model/SpreadMethod/recommend.py:81-111:
        G = getSpreadingGeneralMat(A); W = HybridS(A, G, lambda); F = getResource(A, W);
        recommendForAllUser(F, ...)
    with G cached across lambda values (the findLambda.py:81-116 pattern).
    """

    W_MODES = {"u8x4": (1, 4), "u8x3": (1, 3), "bf16x3": (0, 3), "bf16x2": (0, 2)}

    def __init__(self, n_users: int, n_items: int, users: torch.Tensor, items: torch.Tensor,
                 w_mode: str = "u8x4", g_kind: str = "u8"):
        """w_mode selects how the real operand W of F = A.W reaches the tensor cores:
        "u8x4"/"u8x3": per-column fixed point, 4/3 base-256 digit planes, exact int32 accumulation
                       (|err| <= k_u 2^-(8d+1) s_j; 2 / 1.5 bf16-pass equivalents);
        "bf16x3"/"bf16x2": hi/mid(/lo) bf16 split, fp32 accumulation drained every 8 K-blocks."""
        users = _req(users.to(torch.int32).contiguous(), torch.int32, "users")
        items = _req(items.to(torch.int32).contiguous(), torch.int32, "items")
        if w_mode not in self.W_MODES:
            raise LgcnhsError(f"w_mode must be one of {sorted(self.W_MODES)}")
        self.U, self.M = int(n_users), int(n_items)
        self.dev = users.device
        self.users, self.items = users, items
        self.nnz = int(users.numel())
        self.w_mode = w_mode
        self.w_kind, self.w_planes = self.W_MODES[w_mode]
        self.g_kind = g_kind
        L = lib()
        U, M, dev = self.U, self.M, self.dev
        self.ku = torch.zeros(U, dtype=torch.int32, device=dev)
        self.ki = torch.zeros(M, dtype=torch.int32, device=dev)
        bitmap = torch.zeros((U * M + 31) // 32, dtype=torch.int32, device=dev)
        check(L.hs_degrees(_ptr(users), _ptr(items), self.nnz, U, M, _ptr(self.ku), _ptr(self.ki), _ptr(bitmap),
                           _stream()), "degrees")
        # the deduplicated bit-packed A doubles as the top-k exclusion mask (train ⊕ val items)
        self.excl = ExclusionMask(bitmap, U, M, M)
        # K extent (items) of A and the W^T planes: one 128-byte swizzle atom per K-block
        self.ldM = _pad(M, 128 if self.w_kind == 1 else 64)
        if self.w_kind == 1:
            self.A = torch.zeros((U, self.ldM), dtype=torch.uint8, device=dev)
            check(L.hs_pack_a_u8(_ptr(users), _ptr(items), self.nnz, U, M, _ptr(self.A), self.ldM, _stream()), "pack_a_u8")
            self.col_scale = torch.empty(M, dtype=torch.float32, device=dev)
            self._scratch = torch.empty(20 * M, dtype=torch.uint8, device=dev)
        else:
            self.A = torch.zeros((U, self.ldM), dtype=torch.bfloat16, device=dev)
            check(L.hs_pack_a(_ptr(users), _ptr(items), self.nnz, U, M, _ptr(self.A), self.ldM, _stream()), "pack_a")
            self.col_scale = None
        self.G: Optional[torch.Tensor] = None
        self.C: Optional[torch.Tensor] = None
        self.Wt: Optional[torch.Tensor] = None
        self._topk_scratch: Optional[torch.Tensor] = None

    def excl_items(self, u: int) -> torch.Tensor:
        """Deduplicated item ids of user u (row u of A), ascending — decoded from the bit-packed mask."""
        M = self.M
        b0 = u * M
        w0, w1 = b0 // 32, (b0 + M + 31) // 32
        words = self.excl.bits[w0:w1].to(torch.int64) & 0xFFFFFFFF
        bits = ((words[:, None] >> torch.arange(32, device=words.device)) & 1).flatten()
        return torch.nonzero(bits[b0 - w0 * 32: b0 - w0 * 32 + M]).flatten()

    def fixed_point(self) -> tuple[int, int]:
        """(digits, shift) of the base-256 fixed-point 1/k_u: q_u = round(2^shift / k_u) < 256^digits,
        relative error <= k_max / 2^(shift+1)."""
        nz = self.ku[self.ku > 0]
        kmin, kmax = int(nz.min()), int(nz.max())
        digits = 4
        shift = 8 * digits - 1 + int(math.floor(math.log2(kmin)))
        return digits, shift

    def pack_g_operands(self):
        """(A^T as 0/1 uint8, digit planes Q_d of round(2^shift / k_u) A^T, shift): the K(=user)-major tensor-core
        operands of G, standing in for `A.T / user_degrees` (model/SpreadMethod/model.py:21-25)."""
        U, M, dev = self.U, self.M, self.dev
        ldU = _pad(U, 128)
        digits, shift = self.fixed_point()
        At = torch.zeros((M, ldU), dtype=torch.uint8, device=dev)
        Q = torch.zeros((digits, M, ldU), dtype=torch.uint8, device=dev)
        check(lib().hs_pack_at(_ptr(self.users), _ptr(self.items), self.nnz, U, M, _ptr(self.ku), shift, digits,
                               _ptr(At), _ptr(Q), ldU, M * ldU, _stream()), "pack_at")
        return At, Q, shift

    def general_w(self, item_range: Optional[tuple[int, int]] = None, operands=None,
                  out: Optional[torch.Tensor] = None, symmetric: Optional[bool] = None) -> torch.Tensor:
        """G = A^T K_u^-1 A (model/SpreadMethod/model.py:14-27) on the tensor cores (exact int8 fixed point).
        The full matrix is bitwise symmetric, so by default only the tiles touching its upper triangle are computed
        (hs_gemm_planes_sym); a column block (item_range, the multi-GPU shard) computes all of its tiles."""
        if self.g_kind != "u8":
            raise LgcnhsError("g_kind must be 'u8'")
        j0, j1 = (0, self.M) if item_range is None else item_range
        At, Q, shift = self.pack_g_operands() if operands is None else operands
        if symmetric is None:
            symmetric = item_range is None and os.environ.get("LGCNHS_NO_SYM_G", "0") != "1"
        if symmetric and item_range is not None:
            raise LgcnhsError("general_w: the symmetric schedule needs the full matrix")
        if symmetric:
            M = self.M
            if out is None:
                out = torch.empty((M, _pad(M, 4)), dtype=torch.float32, device=self.dev)[:, :M]
            check(lib().hs_gemm_planes_sym(_ptr(At), int(At.stride(0)), _ptr(Q), int(Q.stride(1)), int(Q.stride(0)),
                                           int(Q.shape[0]), M, self.U, _ptr(out), int(out.stride(0)), 2.0 ** (-shift),
                                           _stream()), "gemm planes (symmetric)")
            G = out
        else:
            G = gemm_planes(1, At, Q[:, j0:j1], self.M, j1 - j0, self.U, out=out, scale=2.0 ** (-shift))
        if item_range is None:
            self.G = G
        return G

    def general_w_allgather(self, group, operands=None, shared=None):
        """Multi-GPU G = A^T K_u^-1 A as a fused GEMM + all-gather (hs_gemm_planes_sym_bcast): the cluster slots of the
        symmetric tile schedule are dealt round-robin to the ranks of `group` (lgcnhs_b200.dist.PeerGroup) and every
        computed tile and its mirror are stored into every rank's G over NVLink — through ONE NVSwitch multicast address
        when the box supports it, else into each peer's replica through CUDA IPC.  `shared` = the value a previous call
        returned (re-uses the shared G buffers).  Returns (G, shared); G is complete on every rank after the group
        barrier this call ends with."""
        At, Q, shift = self.pack_g_operands() if operands is None else operands
        M = self.M
        ldc = _pad(M, 4)
        if shared is None:
            shared = group.shared_matrix(M, ldc)
        G, targets = shared
        arr = (C.c_void_p * len(targets))(*[C.c_void_p(p) for p in targets])
        group.barrier()     # nobody still reads / writes the previous contents of any replica
        check(lib().hs_gemm_planes_sym_bcast(_ptr(At), int(At.stride(0)), _ptr(Q), int(Q.stride(1)), int(Q.stride(0)),
                                             int(Q.shape[0]), M, self.U, arr, len(targets), group.rank, group.world, ldc,
                                             2.0 ** (-shift), _stream()), "gemm planes (symmetric, all-gather)")
        group.barrier()
        self.G = G[:, :M]
        return self.G, shared

    def cooccurrence(self, operands=None) -> torch.Tensor:
        """C = A^T A (common-preference counts, metrics/diversity.py:104) — both operands 0/1, exact int32
        accumulation on the tensor cores; (M, M) fp32 holding integers."""
        if self.C is None:
            At = (self.pack_g_operands()[0] if operands is None else operands[0])
            self.C = gemm_planes(1, At, At.unsqueeze(0), self.M, self.M, self.U)
        return self.C

    def sweep(self, lambdas, k: int, test_pos: Optional[tuple] = None, filtered: bool = True,
              gscore: Optional[torch.Tensor] = None, diversity: bool = True, layer0: Optional[tuple] = None,
              lists_out: Optional[list] = None):
        """The findLambda.py:83-116 loop on the device: G once, then per lambda  HybridS -> A.W [-> * G_score] ->
        filtered top-k -> six metric sums, with no host round trip inside the loop.  The fusion factor G_score
        (getAllocateMat) is either a dense (U, M) matrix `gscore` or, better, `layer0 = (Xu, Xi, seen_csr)`: the
        layer-0 score tiles are then recomputed and multiplied with F inside lgc_score_topk and G_score is never
        materialised.  Returns (sums float64 (n_lambda, 6) on the HOST after ONE device->host copy, list of
        per-lambda metric dicts).  `lists_out` (a list) receives the per-lambda (U, k) id tensors (device) — the
        parity tests check them against the oracle."""
        if self.G is None:
            self.general_w()
        cooc = self.cooccurrence() if diversity else None
        deg = self.ki if diversity else None
        lambdas = [float(x) for x in lambdas]
        sums = torch.zeros((len(lambdas), 6), dtype=torch.float64, device=self.dev)
        fuse = layer0 is None and gscore is None and self.prefer_fused_topk(k, self.U)   # F never materialised
        F = None if fuse else torch.empty((self.U, _pad(self.M, 4)), dtype=torch.float32, device=self.dev)[:, : self.M]
        for n, lam in enumerate(lambdas):
            if layer0 is not None:
                xu, xi, seen = layer0
                self.scale(lam)
                self.resource(out=F)
                idx, _ = score_topk(xu, xi, k, seen, fill=-1024.0, exclude_seen=filtered, mul=F, want_values=False)
            else:
                idx, _ = self.recommend(lam, k, filtered=filtered, gscore=gscore, F_out=F)
            topk_metrics(idx, self.M, test_pos, cooc, deg, out=sums[n])
            if lists_out is not None:
                lists_out.append(idx)
        host = sums.cpu()
        return host, [metrics_from_sums(host[n].tolist(), self.U, k) for n in range(len(lambdas))]

    def scale(self, lam: float, G: Optional[torch.Tensor] = None, want_w32: bool = False):
        """W = HybridS(A, G, lambda) (model/SpreadMethod/model.py:63-85) -> operand planes of W^T (+ fp32 W)."""
        G = self.G if G is None else G
        if G is None:
            raise LgcnhsError("scale(): general_w() has not been computed")
        M, dev = self.M, self.dev
        if self.Wt is None:
            dt = torch.uint8 if self.w_kind == 1 else torch.bfloat16
            self.Wt = torch.zeros((self.w_planes, M, self.ldM), dtype=dt, device=dev)
        W32 = torch.empty((M, M), dtype=torch.float32, device=dev) if want_w32 else None
        if self.w_kind == 1:
            check(lib().hs_scale_w_u8(_ptr(G), int(G.stride(0)), M, _ptr(self.ki), float(lam), _ptr(W32), M,
                                      _ptr(self.Wt), self.ldM, M * self.ldM, self.w_planes, _ptr(self.col_scale),
                                      _ptr(self._scratch), _stream()), "scale_w_u8")
        else:
            check(lib().hs_scale_w(_ptr(G), int(G.stride(0)), M, _ptr(self.ki), float(lam), _ptr(W32), M, _ptr(self.Wt),
                                   self.ldM, M * self.ldM, self.w_planes, _stream()), "scale_w")
        return W32

    def resource(self, user_range: Optional[tuple[int, int]] = None, out: Optional[torch.Tensor] = None):
        """F = A . W (model/SpreadMethod/model.py:88-99) for a block of users."""
        if self.Wt is None:
            raise LgcnhsError("resource(): scale() has not been called")
        u0, u1 = (0, self.U) if user_range is None else user_range
        return gemm_planes(self.w_kind, self.A[u0:u1], self.Wt, u1 - u0, self.M, self.M, out=out, cs=self.col_scale)

    def resource_topk(self, k: int, filtered: bool = True, user_range: Optional[tuple[int, int]] = None,
                      want_values: bool = True):
        """top-k of F = A . W for a block of users WITHOUT materialising F: the selection runs in the epilogue of the
        tensor-core GEMM (hs_resource_topk).  Needs scale() first; k <= 32, more than 128 users, uint8 W planes."""
        if self.Wt is None:
            raise LgcnhsError("resource_topk(): scale() has not been called")
        u0, u1 = (0, self.U) if user_range is None else user_range
        rows = u1 - u0
        nbytes = int(lib().hs_resource_topk_scratch_bytes(rows, self.M, self.w_planes))
        scr = self._topk_scratch
        if scr is None or scr.numel() < nbytes:
            scr = self._topk_scratch = torch.empty(nbytes, dtype=torch.uint8, device=self.dev)
        idx = torch.empty((rows, k), dtype=torch.int64, device=self.dev)
        val = torch.empty((rows, k), dtype=torch.float32, device=self.dev) if want_values else None
        A = self.A[u0:u1]
        check(lib().hs_resource_topk(_ptr(A), int(A.stride(0)), _ptr(self.Wt), int(self.Wt.stride(1)),
                                     int(self.Wt.stride(0)), self.w_planes, rows, self.M, self.M, _ptr(self.col_scale), 1.0,
                                     _ptr(self.excl.bits) if filtered else 0, self.excl.stride_bits if filtered else 0,
                                     u0, int(k), _ptr(idx), _ptr(val), _ptr(scr), nbytes, _stream()), "resource_topk")
        return idx, val

    def can_fuse_topk(self, k: int, rows: int) -> bool:
        return self.w_kind == 1 and k <= 32 and rows > 128

    def prefer_fused_topk(self, k: int, rows: int) -> bool:
        """Default choice of recommend() / sweep().  The fused epilogue never writes F (14.8 GB at the ML-20M shape) but
        its selection runs on the GEMM's four epilogue warps; it pays when F is large (the materialised path is then
        bound by writing and re-reading F) and is requested explicitly (fused=True / LGCNHS_FUSED_TOPK=1) otherwise."""
        if not self.can_fuse_topk(k, rows):
            return False
        env = os.environ.get("LGCNHS_FUSED_TOPK")
        if env is not None:
            return env == "1"
        return rows * self.M * 4 > (2 << 30)

    def recommend(self, lam: float, k: int, filtered: bool = True, gscore: Optional[torch.Tensor] = None,
                  user_range: Optional[tuple[int, int]] = None, F_out: Optional[torch.Tensor] = None,
                  fused: Optional[bool] = None):
        """lambda -> top-k item ids (U, k) int64 + scores; optional fusion F * gscore.
        fused (default: whenever possible and no F_out / gscore is asked for): the top-k is selected inside the
        F-GEMM epilogue and F is never written (resource_topk); otherwise F is materialised and ranked by
        lgc_topk_rows.  Both give identical lists."""
        if self.G is None:
            self.general_w()
        self.scale(lam)
        u0, u1 = (0, self.U) if user_range is None else user_range
        if fused is None:
            fused = gscore is None and F_out is None and self.prefer_fused_topk(k, u1 - u0)
        if fused:
            if gscore is not None:
                raise LgcnhsError("recommend: the fused epilogue has no Hadamard factor; pass fused=False")
            return self.resource_topk(k, filtered, user_range)
        F = self.resource(user_range, out=F_out)
        if gscore is not None:
            check(lib().hs_hadamard(_ptr(F), _ptr(gscore), int(F.shape[0]), int(F.shape[1]), int(F.stride(0)),
                                    int(gscore.stride(0)), _stream()), "hadamard")
        return topk_rows(F, k, self.excl if filtered else None, row_offset=u0)
