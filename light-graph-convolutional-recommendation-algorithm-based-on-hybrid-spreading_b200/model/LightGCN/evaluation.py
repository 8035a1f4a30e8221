"""Drop-in for /root/reference/model/LightGCN/evaluation.py (getValRecommendations, calValLoss)."""
import torch

from lgcnhs_b200 import ops
from lgcnhs_b200.sampling import structured_negative_sampling
from model.LightGCN.model import LightGCN
from utils.graph import convertAdjMatrixToEdgeIndex


_SEEN_CACHE: dict = {}
_EI_CACHE: dict = {}


def _seen_csr_cached(exclude_edge_indices, user_num: int, item_num: int, dev):
    """CSR of the pairs to mask, cached on the identity of the edge tensors: training evaluates with the same train
    edges every epoch_per_eval steps (reference train.py:147-160 rebuilds Python lists of the pairs each time)."""
    key = tuple((e.data_ptr(), tuple(e.shape), e._version) for e in exclude_edge_indices) + (user_num, item_num, str(dev))
    hit = _SEEN_CACHE.get(key)
    if hit is not None and all(a is b for a, b in zip(hit[1], exclude_edge_indices)):
        return hit[0]
    uu = torch.cat([e[0] for e in exclude_edge_indices]).to(dev)
    ii = torch.cat([e[1] for e in exclude_edge_indices]).to(dev)
    seen = ops.seen_csr(uu, ii, user_num, item_num)
    _SEEN_CACHE.clear()
    _SEEN_CACHE[key] = (seen, list(exclude_edge_indices))
    return seen


def _topk_layer0(model, user_num: int, item_num: int, exclude_edge_indices, k: int) -> torch.Tensor:
    """score = e_u^0 . e_i^0^T (LAYER-0 weights, reference evaluation.py:31-34 / recommend.py:83-86),
    seen pairs set to -1024, top-k — one fused kernel, the (U, M) matrix never exists."""
    xu = model.users_emb.weight.detach().contiguous()
    xi = model.items_emb.weight.detach().contiguous()
    dev = xu.device
    seen = _seen_csr_cached(list(exclude_edge_indices), user_num, item_num, dev)
    # one fused kernel: fp32-FMA score tiles -> seen rule -> per-row candidate buffers (lgc_score_topk)
    out, _ = ops.score_topk(xu, xi, k, seen, fill=-float(1 << 10), want_values=False)
    return out


def getValRecommendations(model: LightGCN, user_num: int, item_num: int,
                          train_edge_index: torch.Tensor, val_edge_index: torch.Tensor, k: int) -> torch.Tensor:
    """(U, k) recommendations for validation: only TRAIN pairs are masked (reference evaluation.py:36-52)."""
    key = (train_edge_index.data_ptr(), tuple(train_edge_index.shape), train_edge_index._version, user_num, item_num)
    hit = _EI_CACHE.get(key)
    if hit is not None and hit[1] is train_edge_index:
        train_ei = hit[0]
    else:   # the adjacency is the same tensor at every evaluation of a training run: convert (and mask-CSR) once
        train_ei = convertAdjMatrixToEdgeIndex(user_num, item_num, train_edge_index)
        _EI_CACHE.clear()
        _EI_CACHE[key] = (train_ei, train_edge_index)
    return _topk_layer0(model, user_num, item_num, [train_ei], k)


def calValLoss(model: LightGCN, user_num: int, item_num: int, val_edge_index: torch.Tensor,
               lambda_val: float) -> float:
    """BPR loss over ALL validation edges, propagating over the VAL graph (reference evaluation.py:56-86)."""
    users_f, users_0, items_f, items_0 = model.forward(val_edge_index)
    r_mat = convertAdjMatrixToEdgeIndex(user_num, item_num, val_edge_index)
    u, p, n = structured_negative_sampling(r_mat, contains_neg_self_loops=False)
    # negatives are drawn from [0, max val item id + 1) (lgcnhs_b200/sampling.py): always a valid row of items_f
    E = torch.cat([users_f, items_f]).detach().contiguous()
    X0 = torch.cat([users_0, items_0]).detach().contiguous()
    loss = ops.bpr_fwd_bwd(E, X0, user_num, item_num, u.contiguous(), p.contiguous(), n.contiguous(), lambda_val)
    return round(loss[0].item(), 5)
