"""Negative sampling on the device (the step between forward and loss).

Stands in for torch_geometric.utils.structured_negative_sampling as used at
/root/reference/model/LightGCN/loss.py:58 and evaluation.py:72: for each (u, pos) edge draw
neg ~ U[0, num_nodes) and re-draw while (u, neg) is a positive edge
(contains_neg_self_loops=False additionally forbids neg == u).  The reference moves all E edges
to the host and runs np.isin over them on every training step; here one CUDA thread per requested
triplet (lgc_negative_sample) draws from a counter-based generator and binary-searches the user's
positive row, so only the rows that are needed are touched and nothing leaves the device."""
from __future__ import annotations

from collections import OrderedDict
from typing import Optional

import torch

from ._lib import check, lib
from .ops import seen_csr

_POS_CACHE: "OrderedDict[tuple, tuple]" = OrderedDict()


def _pos_csr(edge_index: torch.Tensor, num_nodes):
    """(num_nodes, sorted positive-item CSR) of the (2, E) user->item edge list, cached on tensor identity — the
    `edge_index.max()` behind PyG's maybe_num_nodes would otherwise cost a device->host sync on every step."""
    key = (edge_index.data_ptr(), tuple(edge_index.shape), edge_index._version, num_nodes)
    hit = _POS_CACHE.get(key)
    if hit is not None and hit[1] is edge_index:
        return hit[0]
    n = int(edge_index.max()) + 1 if num_nodes is None else int(num_nodes)   # max over BOTH rows (quirk P5)
    csr = seen_csr(edge_index[0], edge_index[1], n, n)
    _POS_CACHE[key] = ((n, csr), edge_index)
    while len(_POS_CACHE) > 4:
        _POS_CACHE.popitem(last=False)
    return n, csr


def structured_negative_sampling(edge_index: torch.Tensor, num_nodes: Optional[int] = None,
                                 contains_neg_self_loops: bool = True, rows: Optional[torch.Tensor] = None,
                                 generator: Optional[torch.Generator] = None):
    """Returns (users, pos, neg) for the edges selected by `rows` (all edges if None)."""
    if not edge_index.is_cuda:
        raise RuntimeError("structured_negative_sampling: edge_index must be a CUDA tensor (no CPU fallback)")
    edge_index = edge_index if edge_index.dtype == torch.int64 else edge_index.long()
    dev = edge_index.device
    num_nodes, (ptr, idx) = _pos_csr(edge_index, num_nodes)
    eu, ep = edge_index[0].contiguous(), edge_index[1].contiguous()
    n_edges = int(eu.numel())
    if rows is not None:
        rows = rows.to(device=dev, dtype=torch.int64).contiguous()
    n_out = n_edges if rows is None else int(rows.numel())
    # the seed comes from torch's generator, so torch.manual_seed / an explicit generator make the draw reproducible
    seed = int(torch.randint(0, 2 ** 62, (1,), generator=generator).item())
    out = torch.empty((3, n_out), dtype=torch.int64, device=dev)
    status = torch.zeros(1, dtype=torch.int32, device=dev)
    check(lib().lgc_negative_sample(eu.data_ptr(), ep.data_ptr(), n_edges, 0 if rows is None else rows.data_ptr(), n_out,
                                    ptr.data_ptr(), idx.data_ptr(), num_nodes, num_nodes, int(not contains_neg_self_loops),
                                    seed, out[0].data_ptr(), out[1].data_ptr(), out[2].data_ptr(), status.data_ptr(),
                                    torch.cuda.current_stream().cuda_stream), "negative_sample")
    return out[0], out[1], out[2]
