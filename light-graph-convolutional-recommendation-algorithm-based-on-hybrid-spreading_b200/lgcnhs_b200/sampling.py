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
# status words of the sampler launches since the last check_status(): the kernel flags an out-of-range edge row /
# user id there (and writes the safe triplet (0, 0, 0)); reading them costs a device->host sync, so the training
# loop reads them where it synchronises anyway (the loss read-back at evaluation points)
_PENDING_STATUS: list = []


def _pos_csr(edge_index: torch.Tensor, num_nodes):
    """(num_nodes, sorted positive-item CSR) of the (2, E) user->item edge list, cached on tensor identity — the
    `edge_index.max()` behind PyG's maybe_num_nodes would otherwise cost a device->host sync on every step."""
    key = (edge_index.data_ptr(), tuple(edge_index.shape), edge_index._version, num_nodes)
    hit = _POS_CACHE.get(key)
    if hit is not None and hit[1] is edge_index:
        return hit[0]
    mx = edge_index.max(dim=1).values.tolist()                               # one sync per graph, then cached
    n = max(mx) + 1 if num_nodes is None else int(num_nodes)                 # max over BOTH rows (quirk P5)
    csr = seen_csr(edge_index[0], edge_index[1], n, n)
    _POS_CACHE[key] = ((n, csr, int(mx[1]) + 1), edge_index)
    while len(_POS_CACHE) > 4:
        _POS_CACHE.popitem(last=False)
    return n, csr, int(mx[1]) + 1


def check_status() -> None:
    """Raise if any sampler launch since the last call saw an out-of-range edge row or user id."""
    pending, _PENDING_STATUS[:] = list(_PENDING_STATUS), []
    for st in pending:
        code = int(st.item())
        if code:
            raise RuntimeError(f"lgc_negative_sample: out-of-range {'edge row' if code == 1 else 'user id'} (status {code})")


def structured_negative_sampling(edge_index: torch.Tensor, num_nodes: Optional[int] = None,
                                 contains_neg_self_loops: bool = True, rows: Optional[torch.Tensor] = None,
                                 generator: Optional[torch.Generator] = None, neg_range: Optional[int] = None):
    """Returns (users, pos, neg) for the edges selected by `rows` (all edges if None).

    neg ~ U[0, neg_range).  PyG draws from [0, num_nodes) with num_nodes = max(all ids) + 1; the reference then
    indexes the ITEM table with it (train.py:52-53), which is only in range when the largest id is an item id
    (U <= M: ML-100K, Douban, Amazon-Book) and crashes otherwise (ML-1M, ML-20M).  The default range is therefore
    [0, max item id + 1): identical to the reference wherever the reference runs, and a sound item draw where it
    cannot (no clamping of out-of-range draws onto one item)."""
    if not edge_index.is_cuda:
        raise RuntimeError("structured_negative_sampling: edge_index must be a CUDA tensor (no CPU fallback)")
    edge_index = edge_index if edge_index.dtype == torch.int64 else edge_index.long()
    dev = edge_index.device
    num_nodes, (ptr, idx), n_items_seen = _pos_csr(edge_index, num_nodes)
    neg_range = n_items_seen if neg_range is None else int(neg_range)
    if not 0 < neg_range <= num_nodes:
        raise ValueError(f"neg_range must be in (0, {num_nodes}]")
    eu, ep = edge_index[0].contiguous(), edge_index[1].contiguous()
    n_edges = int(eu.numel())
    if rows is not None:
        rows = rows.to(device=dev, dtype=torch.int64).contiguous()
    n_out = n_edges if rows is None else int(rows.numel())
    # the seed comes from torch's generator, so torch.manual_seed / an explicit generator make the draw reproducible
    seed = int(torch.randint(0, 2 ** 62, (1,), generator=generator).item())
    out = torch.zeros((3, n_out), dtype=torch.int64, device=dev)   # a flagged thread leaves the safe triplet (0, 0, 0)
    status = torch.zeros(1, dtype=torch.int32, device=dev)
    check(lib().lgc_negative_sample(eu.data_ptr(), ep.data_ptr(), n_edges, 0 if rows is None else rows.data_ptr(), n_out,
                                    ptr.data_ptr(), idx.data_ptr(), num_nodes, neg_range, int(not contains_neg_self_loops),
                                    seed, out[0].data_ptr(), out[1].data_ptr(), out[2].data_ptr(), status.data_ptr(),
                                    torch.cuda.current_stream().cuda_stream), "negative_sample")
    _PENDING_STATUS.append(status)
    if len(_PENDING_STATUS) > 64 or rows is None:      # full-edge-list calls (calValLoss) synchronise right after anyway
        check_status()
    return out[0], out[1], out[2]
