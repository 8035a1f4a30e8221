"""TEST INFRASTRUCTURE — independent NumPy evaluation of the reference's six list metrics.

Closed forms of /root/reference/metrics/accurate.py:11-126 and metrics/diversity.py:15-115 (pinned to the
reference's own outputs through tests/golden/metrics_small.npz in test_cpu_host_logic.py).  The GPU tests use
them as the checker of the device kernel lgc_metrics_topk; the product (`metrics/` drop-in) never imports this
file and has no CPU path of its own."""
import numpy as np


def _hits(user_pos_items_dict: dict, rec: np.ndarray):
    """(hit matrix (n_users_in_dict, k), liked counts) in dict iteration order (accurate.py:27-35)."""
    rec = np.asarray(rec, dtype=np.int64)
    uids = np.fromiter(user_pos_items_dict.keys(), dtype=np.int64, count=len(user_pos_items_dict))
    lens = np.fromiter((len(v) for v in user_pos_items_dict.values()), dtype=np.int64, count=uids.size)
    hit = np.zeros((uids.size, rec.shape[1]), dtype=np.float32)
    for r, (u, items) in enumerate(user_pos_items_dict.items()):
        s = set(int(i) for i in items)
        hit[r] = [1.0 if int(i) in s else 0.0 for i in rec[u]]
    return hit, lens.astype(np.float32)


def accurate_metrics(user_pos_items_dict: dict, rec, k: int):
    """(precision, recall, f1, ndcg), each rounded to 5 dp like the reference (fp32 means, accurate.py:36-44,66-95)."""
    hit, liked = _hits(user_pos_items_dict, rec)
    num_correct = hit.sum(-1, dtype=np.float32)
    precision = round(float(np.float32(num_correct.mean(dtype=np.float32)) / k), 5)
    recall = round(float((num_correct / liked).mean(dtype=np.float32)), 5)
    f1 = round(2 * (precision * recall) / (precision + recall), 5) if precision + recall > 0 else float("nan")
    disc = (1.0 / np.log2(np.arange(2, k + 2, dtype=np.float32))).astype(np.float32)
    idcg = np.float32(disc[: min(hit.shape[1], k)].sum(dtype=np.float32))   # ideal = k hits whatever |relevant| is
    dcg = (hit[:, :k] * disc).sum(-1, dtype=np.float32)
    ndcg = round(float((dcg / (idcg if idcg != 0 else 1.0)).mean(dtype=np.float32)), 5)
    return precision, recall, f1, ndcg


def hamming_distance(rec, k: int) -> float:
    """H = 1 - sum_i c_i (c_i - 1) / (U (U-1) k), c_i = number of LISTS containing item i (diversity.py:15-63 intersects
    sets, so an item repeated inside one list counts once)."""
    rec = np.asarray(rec, dtype=np.int64)
    U = rec.shape[0]
    big = int(rec.max()) + 1
    keys = np.unique(np.arange(U)[:, None] * big + rec)
    c = np.bincount(keys % big).astype(np.float64)
    return round(round((U * (U - 1) - np.sum(c * (c - 1)) / k) / (U * (U - 1)), 5), 5)


def internal_similarity(rec, item_degree_dict: dict, interaction_mat: np.ndarray, k: int) -> float:
    """I = sum_u sum_{i != j in L_u} C[i,j] / sqrt(k_i k_j) / (U k (k-1)), C = A^T A (diversity.py:66-115); pairs with a
    zero degree or equal ids are skipped."""
    rec = np.asarray(rec, dtype=np.int64)
    U = rec.shape[0]
    items = np.unique(rec)
    pos = np.searchsorted(items, rec)
    deg = np.array([item_degree_dict.get(int(i), 0) for i in items], dtype=np.float64)
    sub = interaction_mat[:, items]
    C = sub.T @ sub
    with np.errstate(divide="ignore", invalid="ignore"):
        S = C / np.sqrt(np.outer(deg, deg))
    S[~np.isfinite(S)] = 0.0
    S[(deg == 0)[:, None] | (deg == 0)[None, :]] = 0.0
    total = 0.0
    for u in range(U):
        p = pos[u]
        pair = S[p[:, None], p[None, :]]
        total += float(pair[rec[u][:, None] != rec[u][None, :]].sum())
    return round(total / (U * k * (k - 1)), 5)


def diversity_metrics(rec, item_degree_dict: dict, interaction_mat: np.ndarray, k: int):
    return hamming_distance(rec, k), internal_similarity(rec, item_degree_dict, interaction_mat, k)
