"""Fused LightGCN training step shared by model/LightGCN/train.py and model/LightGCNOpti/train.py.

One reference iteration (/root/reference/model/LightGCN/train.py:125-144) is
    forward (gcn_norm + K x gather/mul/scatter_add + stack/mean) -> dense (U+M)^2 round trip ->
    host-side negative sampling over all E edges -> 6 gathers -> ~12 BPR kernels ->
    autograd backward (K x gather/mul/scatter_add again) -> Adam.
Here it is: K fused SpMM launches, one fused BPR forward+backward kernel with the gradient
scatter, K fused SpMM launches on the gradient, two Adam kernels; no host sync inside the step.
"""
from __future__ import annotations

import ctypes as C
import math
import os
from typing import Optional

import torch

from . import ops
from ._lib import check, lib
from .propagation import graphs_for


class FusedBPRTrainer:
    """One reference training iteration as a handful of fused kernels; graph=True (default) captures the whole step —
    2K SpMM layers, BPR, Adam, clean-up — in ONE CUDA graph after two eager warm-up steps: the step is launch-bound on
    small graphs (ML-100K: ~14 launches of a few microseconds each), and every step-dependent scalar (Adam's bias
    corrections, the learning rate) lives in device memory so that the captured graph stays valid.
    LGCNHS_NO_GRAPH=1 forces eager launches.

    Memory traffic of a step beyond the 2K SpMM layers (round 2): the model's two weight tensors are VIEWS of one
    (U+M, D) table, so the forward needs no gather of e^0; the gradient tables are cleaned by zeroing only the rows the
    batch touched (lgc_zero_rows) instead of two full memsets; Adam reads the propagated gradient and the sparse direct
    rows as two operands (lgc_adam_step_fused) instead of a separate add pass.

    distributed=True (BASELINE config 4 at 2/4/8 GPUs; needs an initialised NCCL process group, one
    process per GPU): the rows of A_hat are partitioned over the ranks for BOTH propagations (forward and gradient,
    RowPartitionedPropagation: fused SpMM + peer-store all-gather + device barrier, no NCCL call in the step); every
    rank evaluates the (tiny) BPR kernel on the same mini-batch; Adam runs only on the rows a rank owns and stores the
    updated rows into every rank's weight table over NVLink, followed by one device barrier.  The loss is replicated
    (no reduction needed).  All ranks must draw the same mini-batches (same torch seed).

    Sparse first gradient layer: dL/dE has at most 3 * batch non-zero rows, so the first of the K gradient layers would
    gather ~98 % exact zeros.  The step keeps a one-bit-per-node mask of the batch rows (lgc_row_mask_batch) and the first
    layer skips the masked-out gathers (lgc_propagate_mean_masked / lgc_spmm_rows_bcast_masked): bit-identical gradients.
    Measured on B200 (tools/sparse_backward_bench.py): that layer 101 -> 81 us at the Amazon-Book shape (the (colidx, val)
    stream and the per-row work remain), step 755 -> 745 us; on graphs below ~2 M non-zeros the two mask launches cost what
    the layer saves, so it is used from 2 M non-zeros up (LGCNHS_DENSE_BACKWARD=1 / =0 forces it off / on)."""

    def __init__(self, model, train_adj_index: torch.Tensor, lr: float, eps_reg: float,
                 betas=(0.9, 0.999), adam_eps: float = 1e-8, graph: Optional[bool] = None, distributed: bool = False,
                 deterministic: Optional[bool] = None):
        self.model = model
        self.U, self.M = model.user_num, model.item_num
        self.N, self.D, self.K = self.U + self.M, model.embedding_dim, model.layers
        if not model.users_emb.weight.is_cuda:
            raise RuntimeError("FusedBPRTrainer: the model must live on a CUDA device (no CPU fallback)")
        if self.D not in (32, 64):
            # fail before the first step, not at the first evaluation: the SpMM / BPR / Adam kernels also take 128,
            # the full-rank evaluation kernels (lgc_score_topk, lgc_score_topk_tc) take 32 and 64
            raise RuntimeError(f"FusedBPRTrainer: embedding_dim {self.D} not supported (32 or 64)")
        dev = model.users_emb.weight.device
        self.dev = dev
        self.distributed = bool(distributed)
        # deterministic=True (or LGCNHS_DETERMINISTIC=1): atomic-free gradient scatter -> the whole step is bit-reproducible
        # (the SpMM, the loss reduction and Adam already are); always on in distributed mode, where every rank must hand
        # the gradient propagation the SAME replicated gE
        if deterministic is None:
            deterministic = self.distributed or os.environ.get("LGCNHS_DETERMINISTIC", "0") == "1"
        self.deterministic = bool(deterministic)
        z = lambda: torch.zeros((self.N, self.D), dtype=torch.float32, device=dev)  # noqa: E731
        # one weight table; the module's parameters become views of it (same Parameter objects, same values)
        self.X0 = z()
        with torch.no_grad():
            self.X0[: self.U].copy_(model.users_emb.weight)
            self.X0[self.U:].copy_(model.items_emb.weight)
            model.users_emb.weight.data = self.X0[: self.U]
            model.items_emb.weight.data = self.X0[self.U:]
        self.uw, self.iw = model.users_emb.weight, model.items_emb.weight
        self.gE, self.gX = z(), z()
        self._dense_env = os.environ.get("LGCNHS_DENSE_BACKWARD")
        self.row_mask = ops.row_mask_words(self.N, dev)      # all clear between steps, like gE / gX
        self.exp_avg, self.exp_avg_sq = z(), z()
        self.lr, self.eps_reg, self.betas, self.adam_eps = lr, eps_reg, betas, adam_eps
        self.t = 0
        self.debug_keep_grad = False
        self.grad_total: Optional[torch.Tensor] = None
        self.last_loss: Optional[torch.Tensor] = None
        # device-resident step state (graph mode)
        self.step_dev = torch.zeros(1, dtype=torch.int64, device=dev)
        self.lr_dev = torch.full((1,), float(lr), dtype=torch.float32, device=dev)
        self.hyper_dev = torch.zeros(2, dtype=torch.float32, device=dev)
        self.loss_dev = torch.zeros(2, dtype=torch.float32, device=dev)
        if self.distributed:
            import torch.distributed as dist

            from .dist import RowPartitionedPropagation, _ipc_export, _ipc_import

            if not dist.is_initialized():
                raise RuntimeError("FusedBPRTrainer(distributed=True) needs an initialised process group")
            self.prop = RowPartitionedPropagation(train_adj_index, self.N, self.D, mode="p2p", split=self.U)
            if not torch.equal(self.prop.g.rowptr, self.prop.g.transposed().rowptr):
                raise RuntimeError("distributed training needs the symmetric bipartite adjacency")
            self.rank, self.world = self.prop.rank, self.prop.world
            blobs: list = [None] * self.world
            dist.all_gather_object(blobs, _ipc_export(self.X0))
            self.peer_tables = [self.X0.data_ptr() if r == self.rank else _ipc_import(blobs[r]) for r in range(self.world)]
            self.own_rows = [(a, b) for a, b in self.prop.parts[self.rank] if b > a]
            self.g = self.gt = self.prop.g
            self.E = self.gP = None
            torch.cuda.synchronize()
            dist.barrier()
        else:
            self.g, self.gt = graphs_for(train_adj_index, self.N)
            self.E, self.gP, self.tmp0, self.tmp1 = z(), z(), z(), z()
        self.sparse_backward = (self.g.nnz >= 2_000_000) if self._dense_env is None else self._dense_env != "1"
        if graph is None:
            graph = os.environ.get("LGCNHS_NO_GRAPH", "0") != "1"
        self.use_graph = bool(graph)
        self._graph: Optional[torch.cuda.CUDAGraph] = None
        self._batch: Optional[tuple] = None
        self._eager_steps = 0

    def forward_embeddings(self) -> tuple:
        if self.distributed:
            self.E = self.prop.propagate_mean(self.X0, self.K, result=0)
        else:
            self.g.propagate_mean(self.X0, self.K, out=self.E, tmp=(self.tmp0, self.tmp1))
        return self.X0, self.E

    def _adam(self, a: int, b: int, grad: torch.Tensor) -> None:
        """rows [a, b) of the weight table: g = grad + gX (direct rows), Adam, result into every replica."""
        D = self.D
        if self.distributed:
            arr = (C.c_void_p * self.world)(*[C.c_void_p(p) for p in self.peer_tables])
            n_peers = self.world
        else:
            arr, n_peers = None, 0
        check(lib().lgc_adam_step_fused(self.X0.data_ptr(), arr, n_peers, grad.data_ptr(), self.gX.data_ptr(),
                                        self.exp_avg.data_ptr(), self.exp_avg_sq.data_ptr(), a * D, (b - a) * D,
                                        float(self.betas[0]), float(self.betas[1]), float(self.adam_eps),
                                        self.hyper_dev.data_ptr(), torch.cuda.current_stream().cuda_stream), "adam fused")

    def _body(self, users: torch.Tensor, pos: torch.Tensor, neg: torch.Tensor) -> None:
        """Everything of one step that touches the device; identical in eager and captured form.  gE / gX are all-zero
        on entry and on exit."""
        X0, E = self.forward_embeddings()
        bpr = ops.bpr_fwd_bwd_det if self.deterministic else ops.bpr_fwd_bwd
        bpr(E, X0, self.U, self.M, users, pos, neg, self.eps_reg, self.gE, self.gX, loss_out=self.loss_dev)
        # dL/dX0 = mean_l (A^T)^l dL/dE  +  regulariser rows; dL/dE is non-zero on the batch rows only
        mask = self.row_mask if self.sparse_backward and self.K >= 1 else None
        if mask is not None:
            ops.row_mask_batch(mask, users, pos, neg, self.U, True)
        if self.distributed:
            gP = self.prop.propagate_mean(self.gE, self.K, result=1, x0_row_mask=mask)
        else:
            gP = self.gt.propagate_mean(self.gE, self.K, out=self.gP, tmp=(self.tmp0, self.tmp1), x0_row_mask=mask)
        if mask is not None:
            ops.row_mask_batch(mask, users, pos, neg, self.U, False)
        if self.debug_keep_grad:            # tests only (eager steps): the dense dL/dX0 that Adam consumes as two operands
            self.grad_total = gP + self.gX
        ops.adam_hyper_step(self.step_dev, self.lr_dev, self.betas[0], self.betas[1], self.hyper_dev)
        if self.distributed:
            for a, b in self.own_rows:
                self._adam(a, b, gP)
            self.prop.peer_barrier()        # every rank's new rows are in every replica before the next forward
        else:
            self._adam(0, self.N, gP)
        check(lib().lgc_zero_rows(self.gE.data_ptr(), self.gX.data_ptr(), self.D, users.data_ptr(), pos.data_ptr(),
                                  neg.data_ptr(), int(users.numel()), self.U, torch.cuda.current_stream().cuda_stream),
              "zero rows")

    def _capture(self, batch: int) -> None:
        dev = self.dev
        self._batch = tuple(torch.zeros(batch, dtype=torch.int64, device=dev) for _ in range(3))
        g = torch.cuda.CUDAGraph()
        with torch.no_grad(), torch.cuda.graph(g):
            self._body(*self._batch)
        self._graph = g

    def step(self, users: torch.Tensor, pos: torch.Tensor, neg: torch.Tensor) -> torch.Tensor:
        """One optimisation step on the given triplets; returns the (device) loss tensor [total, bpr]."""
        self.t += 1
        with torch.no_grad():
            if self.use_graph and self._eager_steps >= 2:
                if self._graph is None or self._batch[0].numel() != users.numel():
                    self._capture(int(users.numel()))   # stream capture records the launches, it does not run them
                for dst, src in zip(self._batch, (users, pos, neg)):
                    dst.copy_(src, non_blocking=True)
                self._graph.replay()
            else:
                self._body(users.contiguous(), pos.contiguous(), neg.contiguous())
                self._eager_steps += 1
        self.last_loss = self.loss_dev
        return self.loss_dev

    def decay_lr(self, gamma: float) -> None:
        """ExponentialLR.step() (reference train.py:105,180-181)."""
        self.lr *= gamma
        self.lr_dev.fill_(float(self.lr))

    # algorithmic bytes of one step: 2K SpMM layers + BPR + Adam
    def step_bytes(self, batch: int) -> int:
        return 2 * self.K * self.g.layer_bytes(self.D) + batch * 12 * 4 * self.D + 7 * self.N * 4 * self.D

    # compulsory bytes of one step: every operand moved once (the row gathers of the SpMM are L2-served re-reads)
    def step_bytes_compulsory(self, batch: int) -> int:
        return 2 * self.K * self.g.layer_bytes_compulsory(self.D) + batch * 12 * 4 * self.D + 7 * self.N * 4 * self.D


def user_block(user_num: int, rank: int, world: int) -> tuple:
    """Users [u0, u1) ranked by `rank` when the evaluation is sharded by user block."""
    return user_num * rank // world, user_num * (rank + 1) // world


def gather_user_blocks(idx_block: torch.Tensor, user_num: int, rank: int, world: int) -> torch.Tensor:
    """All-gather of the per-rank (u1 - u0, k) id blocks into the full (user_num, k) list on every rank: the only
    exchange of the sharded evaluation.  Blocks are padded to equal height for one all_gather_into_tensor; works on any
    backend (NCCL on the GPUs, gloo in the CPU tests)."""
    import torch.distributed as dist

    k = int(idx_block.shape[1])
    blk = (user_num + world - 1) // world
    u0, u1 = user_block(user_num, rank, world)
    pad = torch.full((blk, k), -1, dtype=idx_block.dtype, device=idx_block.device)
    pad[: u1 - u0] = idx_block
    out = torch.empty((world * blk, k), dtype=idx_block.dtype, device=idx_block.device)
    dist.all_gather_into_tensor(out, pad)
    rows = torch.cat([torch.arange(b - a, device=idx_block.device) + r * blk
                      for r, (a, b) in enumerate(user_block(user_num, r, world) for r in range(world))])
    return out[rows]


def sharded_topk_layer0(model, user_num: int, item_num: int, seen: tuple, k: int, rank: int, world: int):
    """Full-rank evaluation sharded by USER BLOCK (SURVEY.md 8e "scoring + top-k: independent units"): rank r ranks users
    [U r / P, U (r+1) / P) with the fused score/top-k kernel; the only exchange is the gather of the (U, k) ids."""
    xu = model.users_emb.weight.detach().contiguous()
    xi = model.items_emb.weight.detach().contiguous()
    u0, u1 = user_block(user_num, rank, world)
    idx, _ = ops.score_topk(xu, xi, k, seen, fill=-float(1 << 10), u0=u0, u1=u1, want_values=False)
    return gather_user_blocks(idx, user_num, rank, world), idx


def choose_device() -> torch.device:
    """The reference hard-codes 'cuda:1' (train.py:87); use the current CUDA device and refuse the CPU."""
    if not torch.cuda.is_available():
        raise RuntimeError("trainLightGCN: no CUDA device - the B200 drop-in has no CPU fallback")
    return torch.device("cuda", torch.cuda.current_device())


def safe_diversity(recommendations, item_degree_dict, interaction_mat, k, getDiversityMetrics):
    if interaction_mat is None:
        from metrics.diversity import calHammingDistance

        return calHammingDistance(recommendations, k), math.nan
    return getDiversityMetrics(recommendations, item_degree_dict, interaction_mat, k)
