"""Drop-in for /root/reference/utils/graph.py — the boundary format of LightGCN.forward.

Same results as the reference ((2, 2E) symmetric adjacency COO in the (U+M) node space,
deduplicated, row-major sorted; and its inverse), but without the Python loop over E and the
dense (U+M)^2 fp32 matrix of graph.py:22-31 / :46-48 (83 GB at Amazon-Book shape).  Works on
whatever device the input lives on; sorting/unique is torch plumbing, not a hot-path kernel."""
import torch


def _unique(key: torch.Tensor, bound: int) -> torch.Tensor:
    """Sorted distinct keys.  On the device: the library's own radix sort + compaction (lgc_unique_u64); host tensors
    (the reference's buildGraph hands over CPU tensors) use torch.unique — format plumbing, not a hot-path kernel."""
    if key.is_cuda and key.numel() > 0:
        from lgcnhs_b200 import ops

        return ops.unique_u64(key.contiguous().clone(), bits=max(1, int(bound - 1).bit_length()))
    return torch.unique(key)


def convertEdgeIndexToAdjMatrix(user_num: int, item_num: int, edge_index: torch.Tensor) -> torch.Tensor:
    """(2, E) user->item pairs -> (2, 2E') int64 adjacency indices (reference graph.py:12-35)."""
    u = edge_index[0].to(torch.int64)
    i = edge_index[1].to(torch.int64) + user_num
    n = user_num + item_num
    key = _unique(torch.cat([u * n + i, i * n + u]), n * n)       # dedup + row-major order
    return torch.stack([torch.div(key, n, rounding_mode="floor"), key % n])


def convertAdjMatrixToEdgeIndex(user_num: int, item_num: int, edge_index: torch.Tensor) -> torch.Tensor:
    """(2, 2E) adjacency indices -> (2, E) user->item pairs, row-major sorted (reference graph.py:38-50:
    the [:U, U:] block of the densified matrix)."""
    r, c = edge_index[0].to(torch.int64), edge_index[1].to(torch.int64)
    keep = (r < user_num) & (c >= user_num)
    key = _unique(r[keep] * item_num + (c[keep] - user_num), user_num * item_num)
    return torch.stack([torch.div(key, item_num, rounding_mode="floor"), key % item_num])
