"""ORACLE — TEST INFRASTRUCTURE ONLY.  Never import this from the product path.

CPU restatement (torch, fp32, single thread order of operations kept) of the LightGCN hot
path (P0-P9) of the reference.  Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs may import it.

PARITY STATUS: **unpinned**.  The reference ships no tests, golden vectors or fixtures
(SURVEY.md §4, §8c), and the arithmetic of P1/P2/P5 lives in third-party packages that are
absent from /root/reference and not installable here:
    torch-geometric 2.6.1, torch-scatter 2.1.1, torch-sparse 0.6.17  (environment.yaml:276-278)
Their published algorithms are restated below; the parts that DO live in /root/reference
(model.py's cat / K-loop / stack / mean / split, loss.py, recommend.py) are validated by
oracle/make_golden.py, which imports the reference's real modules on top of oracle/pyg_stub
and checks them against this file (fixtures in tests/golden/).

Each function cites the reference file:line it follows.
"""
from __future__ import annotations

import random
from typing import Optional

import numpy as np
import torch


# --------------------------------------------------------------------------------------------
# P0 / P0'  utils/graph.py
# --------------------------------------------------------------------------------------------
def convert_edge_index_to_adj(user_num: int, item_num: int, edge_index: torch.Tensor) -> torch.Tensor:
    """utils/graph.py:12-35 without the dense (U+M)^2 matrix: the result of
    ``adj_mat.to_sparse_coo().indices()`` is the deduplicated symmetric COO in row-major
    (row, then col) order, users [0,U), items [U,U+M)."""
    u = edge_index[0].to(torch.int64)
    i = edge_index[1].to(torch.int64) + user_num
    n = user_num + item_num
    rows = torch.cat([u, i])
    cols = torch.cat([i, u])
    key = torch.unique(rows * n + cols)  # sorted + dedup == dense scatter then to_sparse_coo
    return torch.stack([torch.div(key, n, rounding_mode="floor"), key % n])


def convert_adj_to_edge_index(user_num: int, item_num: int, adj_index: torch.Tensor) -> torch.Tensor:
    """utils/graph.py:38-50: SparseTensor(row, col).to_dense()[:U, U:].to_sparse_coo().indices()
    == the user->item half, (user, item-U), row-major sorted, deduplicated."""
    r, c = adj_index[0].to(torch.int64), adj_index[1].to(torch.int64)
    keep = (r < user_num) & (c >= user_num)
    key = torch.unique(r[keep] * item_num + (c[keep] - user_num))
    return torch.stack([torch.div(key, item_num, rounding_mode="floor"), key % item_num])


# --------------------------------------------------------------------------------------------
# P1  gcn_norm  (PyG 2.6.1 torch_geometric/nn/conv/gcn_conv.py, dense edge_index branch,
#     add_self_loops=False, edge_weight=None) called at model/LightGCN/model.py:53
# --------------------------------------------------------------------------------------------
def gcn_norm(edge_index: torch.Tensor, num_nodes: Optional[int] = None):
    row, col = edge_index[0], edge_index[1]
    if num_nodes is None:
        num_nodes = int(edge_index.max()) + 1 if edge_index.numel() > 0 else 0  # maybe_num_nodes
    w = torch.ones(edge_index.shape[1], dtype=torch.float32)
    deg = torch.zeros(num_nodes, dtype=torch.float32).scatter_add_(0, col, w)  # scatter(w, col, 'sum')
    dis = deg.pow(-0.5)
    dis.masked_fill_(dis == float("inf"), 0)
    return edge_index, dis[row] * w * dis[col]


# --------------------------------------------------------------------------------------------
# P2  MessagePassing.propagate, aggr="add", flow="source_to_target" + message
#     (model/LightGCN/model.py:61-63, 76-84)
# --------------------------------------------------------------------------------------------
def propagate(edge_index: torch.Tensor, x: torch.Tensor, norm: torch.Tensor) -> torch.Tensor:
    row, col = edge_index[0], edge_index[1]
    x_j = x.index_select(0, row)              # __collect__: x_j = x[edge_index[0]]
    msg = norm.view(-1, 1) * x_j              # message(), model.py:84
    out = torch.zeros_like(x)                 # aggregate(): scatter-add over edge_index[1]
    out.scatter_add_(0, col.view(-1, 1).expand(-1, x.shape[1]), msg)
    return out


def lightgcn_forward(users_w: torch.Tensor, items_w: torch.Tensor, edge_index: torch.Tensor, layers: int):
    """model/LightGCN/model.py:40-74 (identical in LightGCNOpti/model.py:52-86)."""
    ei, norm = gcn_norm(edge_index)                         # :53
    emb_0 = torch.cat([users_w, items_w])                   # :56
    embs = [emb_0]
    emb_k = emb_0
    for _ in range(layers):                                 # :61-63
        emb_k = propagate(ei, emb_k, norm)
        embs.append(emb_k)
    embs = torch.stack(embs, dim=1)                         # :66
    emb_final = torch.mean(embs, dim=1)                     # :69
    users_final, items_final = torch.split(emb_final, [users_w.shape[0], items_w.shape[0]])  # :72
    return users_final, users_w, items_final, items_w      # :74


def propagate_layers(x0: torch.Tensor, edge_index: torch.Tensor, layers: int) -> list:
    """[X0, A X0, ..., A^K X0] — the per-layer tensors of model.py:61-63."""
    ei, norm = gcn_norm(edge_index)
    out = [x0]
    for _ in range(layers):
        out.append(propagate(ei, out[-1], norm))
    return out


# --------------------------------------------------------------------------------------------
# P6  BPRLoss  (model/LightGCN/loss.py:12-44)
# --------------------------------------------------------------------------------------------
def bpr_loss(u_f, u_0, p_f, p_0, n_f, n_0, lambda_val: float) -> torch.Tensor:
    reg = lambda_val * (u_0.norm(2).pow(2) + p_0.norm(2).pow(2) + n_0.norm(2).pow(2))   # :29
    pos = torch.sum(torch.mul(u_f, p_f), dim=-1)                                         # :32-33
    neg = torch.sum(torch.mul(u_f, n_f), dim=-1)                                         # :35-36
    bpr = -torch.mean(torch.nn.functional.softplus(pos - neg))                           # :39 (sic)
    return bpr + reg                                                                     # :42


# --------------------------------------------------------------------------------------------
# P5  structured_negative_sampling (PyG 2.6.1 torch_geometric/utils/_negative_sampling.py)
#     + sampleMiniBatch (model/LightGCN/loss.py:46-70)
# --------------------------------------------------------------------------------------------
def structured_negative_sampling(edge_index: torch.Tensor, num_nodes: Optional[int] = None,
                                 contains_neg_self_loops: bool = True,
                                 generator: Optional[torch.Generator] = None):
    if num_nodes is None:
        num_nodes = int(edge_index.max()) + 1
    row, col = edge_index.cpu()
    pos_idx = row * num_nodes + col
    if not contains_neg_self_loops:
        loop_idx = torch.arange(num_nodes) * (num_nodes + 1)
        pos_idx = torch.cat([pos_idx, loop_idx], dim=0)
    rand = torch.randint(num_nodes, (row.size(0),), dtype=torch.long, generator=generator)
    neg_idx = row * num_nodes + rand
    mask = torch.from_numpy(np.isin(neg_idx.numpy(), pos_idx.numpy())).to(torch.bool)
    rest = mask.nonzero(as_tuple=False).view(-1)
    while rest.numel() > 0:  # rejection: re-draw while (u, neg) is a positive
        tmp = torch.randint(num_nodes, (rest.size(0),), dtype=torch.long, generator=generator)
        rand[rest] = tmp
        neg_idx = row[rest] * num_nodes + tmp
        mask = torch.from_numpy(np.isin(neg_idx.numpy(), pos_idx.numpy())).to(torch.bool)
        rest = rest[mask]
    return edge_index[0], edge_index[1], rand.to(edge_index.device)


def sample_mini_batch(batch_size: int, edge_index: torch.Tensor, generator=None, py_rng: Optional[random.Random] = None):
    edges = torch.stack(structured_negative_sampling(edge_index, generator=generator), dim=0)   # loss.py:58-61
    r = py_rng if py_rng is not None else random
    indices = r.choices([i for i in range(edges[0].shape[0])], k=batch_size)                    # :64 (unseeded)
    batch = edges[:, indices]                                                                   # :67
    return batch[0], batch[1], batch[2]


# --------------------------------------------------------------------------------------------
# P8  recommendForAllUser / getValRecommendations / getAllocateMat
#     (model/LightGCN/recommend.py:83-114; evaluation.py:31-52; SpreadLightGCN/model.py:74-104)
# --------------------------------------------------------------------------------------------
def masked_score(users_w: torch.Tensor, items_w: torch.Tensor, *exclude_edge_indices: torch.Tensor) -> torch.Tensor:
    score = torch.matmul(users_w, items_w.T)          # LAYER-0 weights (recommend.py:83-86)
    for ei in exclude_edge_indices:                   # score[users, items] = -(1 << 10)  (:101, :111)
        score[ei[0], ei[1]] = -(1 << 10)
    return score


def topk_items(score: torch.Tensor, k: int):
    return torch.topk(score, k=k)                     # recommend.py:114


# --------------------------------------------------------------------------------------------
# P7  one training step (model/LightGCN/train.py:125-144) with injected triplets
# --------------------------------------------------------------------------------------------
def train_step(users_w: torch.Tensor, items_w: torch.Tensor, adj_index: torch.Tensor, layers: int, triplets,
               eps: float, optimizer: Optional[torch.optim.Optimizer] = None):
    """users_w/items_w: leaf tensors with requires_grad.  Returns (loss, grad_u, grad_i)."""
    u_f, u_0, i_f, i_0 = lightgcn_forward(users_w, items_w, adj_index, layers)
    u, p, n = triplets
    loss = bpr_loss(u_f[u], u_0[u], i_f[p], i_0[p], i_f[n], i_0[n], eps)   # train.py:55-57, 134-137
    if optimizer is not None:
        optimizer.zero_grad()
    loss.backward()
    gu, gi = users_w.grad.clone(), items_w.grad.clone()
    if optimizer is not None:
        optimizer.step()
    return loss.detach(), gu, gi
