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


class PipelinedPropagation:
    """K-layer propagation of HOST embedding tables with the three legs of a call on three CUDA streams:

        upload e^0 (pinned host -> device)  |  K fused SpMM layers + layer mean  |  download E (device -> pinned host)

    A call returns as soon as its three legs are enqueued; consecutive calls overlap (the upload of call i+1 and the
    download of call i-1 run while call i computes), so the sustained rate is set by the slowest leg instead of their
    sum.  Two device buffer sets alternate between calls.  `synchronize()` waits for everything in flight; the result of
    call i is valid in its `out_host` after `synchronize()` or after the event `submit` returns has completed.

    This is the serving-style entry point behind LightGCN.forward for callers whose embeddings live on the host
    (model/LightGCN/model.py:40-74 computes the same E; the reference moves the whole module to the device instead)."""

    def __init__(self, edge_index: torch.Tensor, n_nodes: int, dim: int, layers: int, depth: int = 2):
        if not edge_index.is_cuda:
            raise RuntimeError("PipelinedPropagation: edge_index must be a CUDA tensor (no CPU fallback)")
        self.g, _ = graphs_for(edge_index, n_nodes)
        self.layers = layers
        dev = edge_index.device
        z = lambda: torch.empty((n_nodes, dim), dtype=torch.float32, device=dev)  # noqa: E731
        self.slots = [{"x0": z(), "e": z(), "tmp": (z(), z()), "up": torch.cuda.Event(), "done": torch.cuda.Event(),
                       "down": torch.cuda.Event()} for _ in range(depth)]
        self.s_up, self.s_cmp, self.s_down = (torch.cuda.Stream(device=dev) for _ in range(3))
        self.calls = 0

    def submit(self, x0_host: torch.Tensor, out_host: torch.Tensor) -> torch.cuda.Event:
        """Enqueue upload -> propagate -> download for one (n_nodes, dim) pinned host table; returns the event that
        completes when `out_host` holds E."""
        if not (x0_host.is_pinned() and out_host.is_pinned()):
            raise RuntimeError("PipelinedPropagation: host tables must be pinned (torch.Tensor.pin_memory())")
        sl = self.slots[self.calls % len(self.slots)]
        self.calls += 1
        with torch.cuda.stream(self.s_up):
            self.s_up.wait_event(sl["done"])          # the previous propagation that read this slot's x0 has finished
            sl["x0"].copy_(x0_host, non_blocking=True)
            sl["up"].record(self.s_up)
        with torch.cuda.stream(self.s_cmp):
            self.s_cmp.wait_event(sl["up"])
            self.s_cmp.wait_event(sl["down"])         # the previous download of this slot's E has finished
            self.g.propagate_mean(sl["x0"], self.layers, out=sl["e"], tmp=sl["tmp"])
            sl["done"].record(self.s_cmp)
        with torch.cuda.stream(self.s_down):
            self.s_down.wait_event(sl["done"])
            out_host.copy_(sl["e"], non_blocking=True)
            sl["down"].record(self.s_down)
        return sl["down"]

    def synchronize(self) -> None:
        for s in (self.s_up, self.s_cmp, self.s_down):
            s.synchronize()
