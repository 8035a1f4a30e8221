"""Autograd glue for the fused K-layer propagation + the per-graph CSR cache.

Stands in for the body of LightGCN.forward (/root/reference/model/LightGCN/model.py:53-72):
gcn_norm on every call + K x (index_select, mul, scatter_add) + stack + mean.  Here the
normalised CSR is built once per edge_index tensor and the K layers run as K fused SpMM
launches (Horner form of the layer mean); the backward pass is the same kernel on the
transposed graph (identical arrays when the graph is symmetric, as the bipartite adjacency is).
"""
from __future__ import annotations

from collections import OrderedDict

import torch

from .ops import NormGraph

_CACHE: "OrderedDict[tuple, tuple]" = OrderedDict()
_CACHE_MAX = 6


def graphs_for(edge_index: torch.Tensor, n_nodes: int):
    """(forward graph, backward graph) for an edge_index tensor, cached by tensor identity + version."""
    key = (edge_index.data_ptr(), tuple(edge_index.shape), edge_index._version, str(edge_index.device), n_nodes)
    hit = _CACHE.get(key)
    if hit is not None and hit[2] is edge_index:
        _CACHE.move_to_end(key)
        return hit[0], hit[1]
    ei = edge_index.contiguous()
    g = NormGraph(ei, n_nodes)
    # backward graph = structural transpose with g's own values (NOT gcn_norm of the transposed edge list, which
    # would normalise by out-degrees); for the symmetric bipartite adjacency it is g itself
    gt = g.transposed()
    if torch.equal(g.rowptr, gt.rowptr) and torch.equal(g.colidx, gt.colidx) and torch.equal(g.val, gt.val):
        gt = g  # symmetric graph: A_hat^T == A_hat, share the arrays
    _CACHE[key] = (g, gt, edge_index)  # keeps the tensor alive so data_ptr cannot be recycled
    while len(_CACHE) > _CACHE_MAX:
        _CACHE.popitem(last=False)
    return g, gt


def clear_cache() -> None:
    _CACHE.clear()


class PropagateMean(torch.autograd.Function):
    """E = mean_{l=0..K} A_hat^l X0;  dX0 = mean_{l=0..K} (A_hat^T)^l dE."""

    @staticmethod
    def forward(ctx, x0: torch.Tensor, g: NormGraph, gt: NormGraph, layers: int):
        ctx.gt, ctx.layers = gt, layers
        return g.propagate_mean(x0.contiguous(), layers)

    @staticmethod
    def backward(ctx, grad_out: torch.Tensor):
        return ctx.gt.propagate_mean(grad_out.contiguous(), ctx.layers), None, None, None


def lightgcn_forward(users_w: torch.Tensor, items_w: torch.Tensor, edge_index: torch.Tensor, layers: int):
    """The 4-tuple of LightGCN.forward (model.py:74) on the CUDA path.  No CPU fallback."""
    if not (users_w.is_cuda and items_w.is_cuda and edge_index.is_cuda):
        raise RuntimeError("LightGCN.forward: weights and edge_index must be CUDA tensors - "
                           "the B200 drop-in has no CPU fallback (move the model with .to('cuda'))")
    n_users, n_items = users_w.shape[0], items_w.shape[0]
    g, gt = graphs_for(edge_index, n_users + n_items)
    emb_0 = torch.cat([users_w, items_w])
    emb_final = PropagateMean.apply(emb_0, g, gt, layers)
    users_final, items_final = torch.split(emb_final, [n_users, n_items])
    return users_final, users_w, items_final, items_w
