"""Drop-in for /root/reference/metrics/diversity.py — Hamming distance H and intra-list
similarity I of the top-k lists, as closed forms instead of the reference's O(U^2) Python pair
loop with string-keyed memo (diversity.py:15-63) and O(U k^2 U) dot products (:66-115):

    H = 1 - sum_i c_i (c_i - 1) / (U (U-1) k)       c_i = number of lists containing item i
    I = sum_u sum_{i != j in L_u} C[i,j] / sqrt(k_i k_j) / (U k (k-1)),   C = A^T A

(both identities are pinned to the reference's outputs at its 5-dp rounding by tests/golden/metrics_small.npz).
They run on the device: item histogram + exact int8 tensor-core co-occurrence GEMM + one warp per
user (lgc_metrics_topk).  No CPU path."""
import numpy as np
import torch

from utils.log import logger
from utils.wrapper import calTimes


def _require_cuda() -> None:
    if not torch.cuda.is_available():
        raise RuntimeError("metrics.diversity: no CUDA device - the B200 drop-in has no CPU fallback")


@calTimes(logger, "海明距离计算完成")
def calHammingDistance(recommendations: torch.Tensor, k: int) -> float:
    """reference diversity.py:15-63."""
    _require_cuda()
    from lgcnhs_b200.metrics_device import hamming_device
    return hamming_device(recommendations, k)


@calTimes(logger, "内部相似性计算完成")
def calInternalSimilarity(recommendations: torch.Tensor, item_degree_dict: dict,
                          interaction_mat: np.ndarray, k: int) -> float:
    """reference diversity.py:66-115."""
    _require_cuda()
    from lgcnhs_b200.metrics_device import diversity_device
    return diversity_device(recommendations, item_degree_dict, interaction_mat, k)[1]


def getDiversityMetrics(recommendations: torch.Tensor, item_degree_dict: dict,
                        interaction_mat: np.ndarray, k: int) -> tuple:
    """(H, I) — reference diversity.py:118-135."""
    _require_cuda()
    from lgcnhs_b200.metrics_device import diversity_device
    return diversity_device(recommendations, item_degree_dict, interaction_mat, k)
