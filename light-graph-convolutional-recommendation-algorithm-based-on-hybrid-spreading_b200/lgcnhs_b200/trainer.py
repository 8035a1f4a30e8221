"""Fused LightGCN training step shared by model/LightGCN/train.py and model/LightGCNOpti/train.py.

One reference iteration (/root/reference/model/LightGCN/train.py:125-144) is
    forward (gcn_norm + K x gather/mul/scatter_add + stack/mean) -> dense (U+M)^2 round trip ->
    host-side negative sampling over all E edges -> 6 gathers -> ~12 BPR kernels ->
    autograd backward (K x gather/mul/scatter_add again) -> Adam.
Here it is: K fused SpMM launches, one fused BPR forward+backward kernel with the gradient
scatter, K fused SpMM launches on the gradient, two Adam kernels; no host sync inside the step.
"""
from __future__ import annotations

import math
import os
from typing import Optional

import torch

from . import ops
from .propagation import graphs_for


class FusedBPRTrainer:
    """graph=True (default) captures the whole step — 2K SpMM layers, BPR, gradient merge, Adam — in ONE CUDA graph
    after two eager warm-up steps: the step is launch-bound on small graphs (ML-100K: ~14 launches of a few
    microseconds each), and every step-dependent scalar (Adam's bias corrections, the learning rate) lives in
    device memory so that the captured graph stays valid.  Set LGCNHS_NO_GRAPH=1 to force eager launches."""

    def __init__(self, model, train_adj_index: torch.Tensor, lr: float, eps_reg: float,
                 betas=(0.9, 0.999), adam_eps: float = 1e-8, graph: Optional[bool] = None):
        self.model = model
        self.U, self.M = model.user_num, model.item_num
        self.N, self.D, self.K = self.U + self.M, model.embedding_dim, model.layers
        self.uw, self.iw = model.users_emb.weight, model.items_emb.weight
        if not self.uw.is_cuda:
            raise RuntimeError("FusedBPRTrainer: the model must live on a CUDA device (no CPU fallback)")
        if self.D not in (32, 64):
            # fail before the first step, not at the first evaluation: the SpMM / BPR / Adam kernels also take 128,
            # the full-rank evaluation kernel (lgc_score_topk) takes 32 and 64
            raise RuntimeError(f"FusedBPRTrainer: embedding_dim {self.D} not supported (32 or 64)")
        dev = self.uw.device
        self.dev = dev
        self.g, self.gt = graphs_for(train_adj_index, self.N)
        z = lambda: torch.zeros((self.N, self.D), dtype=torch.float32, device=dev)  # noqa: E731
        self.X0, self.E, self.gE, self.gX, self.tmp0, self.tmp1, self.gP = z(), z(), z(), z(), z(), z(), z()
        self.exp_avg, self.exp_avg_sq = z(), z()
        self.lr, self.eps_reg, self.betas, self.adam_eps = lr, eps_reg, betas, adam_eps
        self.t = 0
        self.last_loss: Optional[torch.Tensor] = None
        # device-resident step state (graph mode)
        self.step_dev = torch.zeros(1, dtype=torch.int64, device=dev)
        self.lr_dev = torch.full((1,), float(lr), dtype=torch.float32, device=dev)
        self.hyper_dev = torch.zeros(2, dtype=torch.float32, device=dev)
        self.loss_dev = torch.zeros(2, dtype=torch.float32, device=dev)
        if graph is None:
            graph = os.environ.get("LGCNHS_NO_GRAPH", "0") != "1"
        self.use_graph = bool(graph)
        self._graph: Optional[torch.cuda.CUDAGraph] = None
        self._batch: Optional[tuple] = None
        self._eager_steps = 0

    def forward_embeddings(self) -> tuple:
        U = self.U
        self.X0[:U].copy_(self.uw.detach())
        self.X0[U:].copy_(self.iw.detach())
        self.g.propagate_mean(self.X0, self.K, out=self.E, tmp=(self.tmp0, self.tmp1))
        return self.X0, self.E

    def _body(self, users: torch.Tensor, pos: torch.Tensor, neg: torch.Tensor) -> None:
        """Everything of one step that touches the device; identical in eager and captured form."""
        X0, E = self.forward_embeddings()
        self.gE.zero_()
        self.gX.zero_()
        loss = ops.bpr_fwd_bwd(E, X0, self.U, self.M, users, pos, neg, self.eps_reg, self.gE, self.gX)
        self.loss_dev.copy_(loss)
        # dL/dX0 = mean_l (A^T)^l dL/dE  +  regulariser rows
        self.gt.propagate_mean(self.gE, self.K, out=self.gP, tmp=(self.tmp0, self.tmp1))
        self.gX.add_(self.gP)
        U = self.U
        ops.adam_hyper_step(self.step_dev, self.lr_dev, self.betas[0], self.betas[1], self.hyper_dev)
        ops.adam_step_dev(self.uw.data, self.gX[:U], self.exp_avg[:U], self.exp_avg_sq[:U], self.betas[0], self.betas[1],
                          self.adam_eps, self.hyper_dev)
        ops.adam_step_dev(self.iw.data, self.gX[U:], self.exp_avg[U:], self.exp_avg_sq[U:], self.betas[0], self.betas[1],
                          self.adam_eps, self.hyper_dev)

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
