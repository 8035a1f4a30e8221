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
from typing import Optional

import torch

from . import ops
from .propagation import graphs_for


class FusedBPRTrainer:
    def __init__(self, model, train_adj_index: torch.Tensor, lr: float, eps_reg: float,
                 betas=(0.9, 0.999), adam_eps: float = 1e-8):
        self.model = model
        self.U, self.M = model.user_num, model.item_num
        self.N, self.D, self.K = self.U + self.M, model.embedding_dim, model.layers
        self.uw, self.iw = model.users_emb.weight, model.items_emb.weight
        if not self.uw.is_cuda:
            raise RuntimeError("FusedBPRTrainer: the model must live on a CUDA device (no CPU fallback)")
        dev = self.uw.device
        self.g, self.gt = graphs_for(train_adj_index, self.N)
        z = lambda: torch.zeros((self.N, self.D), dtype=torch.float32, device=dev)  # noqa: E731
        self.E, self.gE, self.gX, self.tmp0, self.tmp1, self.gP = z(), z(), z(), z(), z(), z()
        self.exp_avg, self.exp_avg_sq = z(), z()
        self.lr, self.eps_reg, self.betas, self.adam_eps = lr, eps_reg, betas, adam_eps
        self.t = 0
        self.last_loss: Optional[torch.Tensor] = None

    def forward_embeddings(self) -> tuple:
        X0 = torch.cat([self.uw.detach(), self.iw.detach()])
        self.g.propagate_mean(X0, self.K, out=self.E, tmp=(self.tmp0, self.tmp1))
        return X0, self.E

    def step(self, users: torch.Tensor, pos: torch.Tensor, neg: torch.Tensor) -> torch.Tensor:
        """One optimisation step on the given triplets; returns the (device) loss tensor [total, bpr]."""
        X0, E = self.forward_embeddings()
        self.gE.zero_()
        self.gX.zero_()
        loss = ops.bpr_fwd_bwd(E, X0, self.U, self.M, users, pos, neg, self.eps_reg, self.gE, self.gX)
        # dL/dX0 = mean_l (A^T)^l dL/dE  +  regulariser rows
        self.gt.propagate_mean(self.gE, self.K, out=self.gP, tmp=(self.tmp0, self.tmp1))
        self.gX.add_(self.gP)
        self.t += 1
        U = self.U
        with torch.no_grad():
            ops.adam_step(self.uw.data, self.gX[:U], self.exp_avg[:U], self.exp_avg_sq[:U], self.t, self.lr,
                          self.betas[0], self.betas[1], self.adam_eps)
            ops.adam_step(self.iw.data, self.gX[U:], self.exp_avg[U:], self.exp_avg_sq[U:], self.t, self.lr,
                          self.betas[0], self.betas[1], self.adam_eps)
        self.last_loss = loss
        return loss

    def decay_lr(self, gamma: float) -> None:
        """ExponentialLR.step() (reference train.py:105,180-181)."""
        self.lr *= gamma

    # algorithmic bytes of one step: 2K SpMM layers + BPR + Adam
    def step_bytes(self, batch: int) -> int:
        return 2 * self.K * self.g.layer_bytes(self.D) + batch * 12 * 4 * self.D + 7 * self.N * 4 * self.D


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
