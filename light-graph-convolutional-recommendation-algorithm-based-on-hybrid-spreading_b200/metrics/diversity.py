"""Drop-in for /root/reference/metrics/diversity.py — Hamming distance H and intra-list
similarity I of the top-k lists, as closed forms instead of the reference's O(U^2) Python pair
loop with string-keyed memo (diversity.py:15-63) and O(U k^2 U) dot products (:66-115):

    H = 1 - sum_i c_i (c_i - 1) / (U (U-1) k)       c_i = number of lists containing item i
    I = sum_u sum_{i != j in L_u} C[i,j] / sqrt(k_i k_j) / (U k (k-1)),   C = A^T A

(both identities checked against the reference's code to its 5-dp rounding)."""
import numpy as np
import torch

from utils.log import logger
from utils.wrapper import calTimes


@calTimes(logger, "海明距离计算完成")
def calHammingDistance(recommendations: torch.Tensor, k: int) -> float:
    rec = recommendations.detach().cpu().numpy().astype(np.int64)
    user_num = rec.shape[0]
    # the reference intersects *sets*, so an item repeated inside one list counts once
    keys = np.unique(np.arange(user_num)[:, None] * (rec.max() + 1) + rec)
    c = np.bincount(keys % (rec.max() + 1)).astype(np.float64)
    shared = np.sum(c * (c - 1))                     # ordered user pairs x common items
    total_h = user_num * (user_num - 1) - shared / k
    return round(round(total_h / (user_num * (user_num - 1)), 5), 5)


@calTimes(logger, "内部相似性计算完成")
def calInternalSimilarity(recommendations: torch.Tensor, item_degree_dict: dict,
                          interaction_mat: np.ndarray, k: int) -> float:
    rec = recommendations.detach().cpu().numpy().astype(np.int64)
    user_num, width = rec.shape
    items = np.unique(rec)
    pos = np.searchsorted(items, rec)                # list entries as indices into `items`
    deg = np.array([item_degree_dict.get(int(i), 0) for i in items], dtype=np.float64)
    sub = interaction_mat[:, items]
    C = sub.T @ sub                                  # common-preference counts of the recommended items
    with np.errstate(divide="ignore", invalid="ignore"):
        S = C / np.sqrt(np.outer(deg, deg))
    S[~np.isfinite(S)] = 0.0                         # k_i == 0 or k_j == 0 pairs are skipped
    S[np.outer(deg == 0, np.ones_like(deg, dtype=bool)) | np.outer(np.ones_like(deg, dtype=bool), deg == 0)] = 0.0
    total = 0.0
    blk = max(1, (1 << 24) // max(1, width * width))
    for s in range(0, user_num, blk):
        p = pos[s:s + blk]
        pair = S[p[:, :, None], p[:, None, :]]       # (b, k, k)
        same = rec[s:s + blk][:, :, None] == rec[s:s + blk][:, None, :]   # iid_i == iid_j pairs are skipped
        total += float(pair[~same].sum())
    return round(total / (user_num * k * (k - 1)), 5)


def getDiversityMetrics(recommendations: torch.Tensor, item_degree_dict: dict,
                        interaction_mat: np.ndarray, k: int) -> tuple:
    if torch.cuda.is_available():
        # device path (lgc_metrics_topk): item histogram + exact int8 tensor-core co-occurrence GEMM
        from lgcnhs_b200.metrics_device import diversity_device
        return diversity_device(recommendations, item_degree_dict, interaction_mat, k)
    H = calHammingDistance(recommendations, k)
    I = calInternalSimilarity(recommendations, item_degree_dict, interaction_mat, k)
    return H, I
