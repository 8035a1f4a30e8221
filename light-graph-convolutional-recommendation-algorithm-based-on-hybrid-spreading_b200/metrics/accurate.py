"""Drop-in for /root/reference/metrics/accurate.py — Precision / Recall / F1 / NDCG of top-k lists.
Same functions, arguments, rounding (5 dp) and quirks (NDCG's ideal list is k hits regardless of
how many relevant items a user has, accurate.py:76-86).  The per-user Python `map(lambda ...)`
membership loops of the reference run as one device kernel (lgc_metrics_topk, one warp per user:
binary search of every recommended id in the user's relevant-item row).  No CPU path."""
import torch


def _device_metrics(user_pos_items_dict: dict, recommendations: torch.Tensor, k: int) -> tuple:
    if not torch.cuda.is_available():
        raise RuntimeError("metrics.accurate: no CUDA device - the B200 drop-in has no CPU fallback")
    if not len(user_pos_items_dict):
        raise ValueError("metrics.accurate: empty user_pos_items_dict (the reference divides by its length)")
    from lgcnhs_b200.metrics_device import accuracy_device
    return accuracy_device(user_pos_items_dict, recommendations, k)


def calPrecisionAndRecall(user_pos_items_dict: dict, recommendations: torch.Tensor, k: int) -> tuple:
    """reference accurate.py:11-47."""
    p, r, _, _ = _device_metrics(user_pos_items_dict, recommendations, k)
    return p, r


def calF1Score(precision: float, recall: float) -> float:
    """reference accurate.py:49-63 (host arithmetic on two scalars)."""
    return round(2 * (precision * recall) / (precision + recall), 5)


def calNDCG(user_pos_items_dict: dict, recommendations: torch.Tensor, k: int) -> float:
    """reference accurate.py:66-98."""
    return _device_metrics(user_pos_items_dict, recommendations, k)[3]


def getAccurateMetrics(user_pos_items_dict: dict, recommendations: torch.Tensor, k: int) -> tuple:
    """(precision, recall, f1, ndcg) — reference accurate.py:101-126, one kernel for all four."""
    return _device_metrics(user_pos_items_dict, recommendations, k)
