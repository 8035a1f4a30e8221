"""Drop-in for /root/reference/metrics/accurate.py — Precision / Recall / F1 / NDCG of top-k lists.
Same functions, arguments, rounding (5 dp) and quirks (NDCG's ideal list is k hits regardless of
how many relevant items a user has, accurate.py:76-86); the per-user Python `map(lambda ...)`
membership loops become one vectorised sorted-key lookup."""
import numpy as np
import torch


def _hits(user_pos_items_dict: dict, recommendations: torch.Tensor) -> tuple:
    """(hit matrix (n_users_in_dict, k) float32, liked counts) in dict iteration order."""
    rec = recommendations.detach().cpu().numpy().astype(np.int64)
    uids = np.fromiter(user_pos_items_dict.keys(), dtype=np.int64, count=len(user_pos_items_dict))
    lens = np.fromiter((len(v) for v in user_pos_items_dict.values()), dtype=np.int64, count=uids.size)
    big = int(max(rec.max(initial=0), max((max(v) for v in user_pos_items_dict.values() if len(v)), default=0))) + 1
    pos_keys = np.concatenate([np.asarray(v, dtype=np.int64) for v in user_pos_items_dict.values()] or [np.empty(0, np.int64)])
    pos_keys = np.unique(np.repeat(uids, lens) * big + pos_keys)
    rec_keys = uids[:, None] * big + rec[uids]
    pos = np.searchsorted(pos_keys, rec_keys)
    pos[pos == pos_keys.size] = 0
    hit = (pos_keys[pos] == rec_keys) if pos_keys.size else np.zeros_like(rec_keys, dtype=bool)
    return torch.from_numpy(hit.astype(np.float32)), torch.from_numpy(lens.astype(np.float32))


def calPrecisionAndRecall(user_pos_items_dict: dict, recommendations: torch.Tensor, k: int) -> tuple:
    hit, liked = _hits(user_pos_items_dict, recommendations)
    num_correct = torch.sum(hit, dim=-1)
    precision = torch.mean(num_correct) / k
    recall = torch.mean(num_correct / liked)
    return round(precision.item(), 5), round(recall.item(), 5)


def calF1Score(precision: float, recall: float) -> float:
    return round(2 * (precision * recall) / (precision + recall), 5)


def calNDCG(user_pos_items_dict: dict, recommendations: torch.Tensor, k: int) -> float:
    hit, _ = _hits(user_pos_items_dict, recommendations)
    disc = 1. / torch.log2(torch.arange(2, k + 2))
    length = min(hit.shape[1], k)
    ideal = torch.zeros((hit.shape[0], k))
    ideal[:, :length] = 1                      # reference: all recommended positions count as relevant
    idcg = torch.sum(ideal * disc, axis=1)
    dcg = torch.sum(hit * disc, axis=1)
    idcg[idcg == 0.] = 1.
    ndcg = dcg / idcg
    ndcg[torch.isnan(ndcg)] = 0.
    return round(torch.mean(ndcg).item(), 5)


def getAccurateMetrics(user_pos_items_dict: dict, recommendations: torch.Tensor, k: int) -> tuple:
    if torch.cuda.is_available() and len(user_pos_items_dict):
        from lgcnhs_b200.metrics_device import accuracy_device   # lgc_metrics_topk: one warp per user
        return accuracy_device(user_pos_items_dict, recommendations, k)
    precision, recall = calPrecisionAndRecall(user_pos_items_dict, recommendations, k)
    f1 = calF1Score(precision, recall)
    ndcg = calNDCG(user_pos_items_dict, recommendations, k)
    return precision, recall, f1, ndcg
