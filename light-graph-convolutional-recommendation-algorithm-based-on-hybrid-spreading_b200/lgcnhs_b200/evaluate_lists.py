"""The reference's multi-k evaluation (/root/reference/evaluationMetrics.py:43-96) as batched device calls.

evaluationMetrics.py loops over k in {30, 50, 100} and six model names; for EVERY (k, model) it re-reads the saved
recommendation dict, rebuilds the positive-item dicts, the item degrees and the dense interaction matrix of train + val with
Python loops, and then runs the O(U^2) / O(U k^2 U) metric loops.  Here the data-set side is built ONCE on the device — the
test-positive CSR, the item degrees and the co-occurrence matrix C = A^T A (exact int8 tensor-core GEMM) — every list set
is one lgc_metrics_topk launch, and the whole (n_sets, 6) table of sums crosses PCIe once.  The result has the reference's
per-k sheets (columns Model, P, R, F1, NDCG, H, I; evaluationMetrics.py:74-82)."""
from __future__ import annotations

from typing import Mapping, Optional, Sequence, Union

import numpy as np
import pandas as pd
import torch

from . import ops
from .recommend_common import cuda_device, interactions_from_frames

Lists = Union[Mapping[int, Sequence[int]], np.ndarray, torch.Tensor]


def lists_to_tensor(lists: Lists, user_num: int, k: int, dev: torch.device) -> torch.Tensor:
    """dict{uid: ids} (the saved .npy format, recommend.py:122) or a (U, >= k) array -> (U, k) int64 on the device, users in
    id order like recommendDictToTensor (utils/trans.py)."""
    if isinstance(lists, torch.Tensor):
        rec = lists.detach().to(dev).long()
    elif isinstance(lists, np.ndarray):
        rec = torch.from_numpy(np.ascontiguousarray(lists)).to(dev).long()
    else:
        rows = [np.asarray(lists[u], dtype=np.int64)[:k] for u in range(user_num)]
        if any(r.size < k for r in rows):
            raise ValueError(f"evaluate_lists: a recommendation list is shorter than k={k}")
        rec = torch.from_numpy(np.stack(rows)).to(dev)
    if rec.dim() != 2 or rec.shape[0] != user_num or rec.shape[1] < k:
        raise ValueError(f"evaluate_lists: expected {user_num} lists of at least {k} ids, got {tuple(rec.shape)}")
    return rec[:, :k].contiguous()


def pos_csr_keep_duplicates(users: torch.Tensor, items: torch.Tensor, user_num: int, dev: torch.device) -> tuple:
    """(user, item) rows -> int32 CSR over all users, items ascending per row, duplicates KEPT (the reference divides by
    len(items), accurate.py:33): own radix sort of the (user, item) keys, row pointer from the sorted user ids."""
    users, items = users.to(dev).long(), items.to(dev).long()
    if users.numel() == 0:
        return torch.zeros(user_num + 1, dtype=torch.int32, device=dev), torch.zeros(1, dtype=torch.int32, device=dev)
    stride = int(items.max()) + 1
    keys = ops.sort_u64((users * stride + items).contiguous(), bits=max(1, (user_num * stride - 1).bit_length()))
    rows = keys // stride
    ptr = torch.searchsorted(rows, torch.arange(user_num + 1, device=dev, dtype=torch.int64)).to(torch.int32)
    return ptr, (keys % stride).to(torch.int32).contiguous()


def evaluate_lists(user_num: int, item_num: int, train_data_df: pd.DataFrame, val_data_df: pd.DataFrame,
                   test_data_df: pd.DataFrame, lists: Mapping[tuple, Lists], save_path: Optional[str] = None) -> dict:
    """lists[(model_name, k)] = the recommendation lists of that model at length k.  Returns {k: DataFrame(Model, P, R, F1,
    NDCG, H, I)} in the insertion order of `lists`; with save_path the sheets are also written as
    `model_evaluation_results_<k>.csv` (the reference writes one .xlsx with a sheet per k, which needs openpyxl)."""
    dev = cuda_device()
    u, i = interactions_from_frames(train_data_df, val_data_df)
    eng = ops.SpreadingEngine(user_num, item_num, u, i)
    cooc = eng.cooccurrence()
    # item degree = occurrences in the train + val positive lists (getItemDegreeByUserPosItemDict), duplicates counted
    deg = torch.bincount(i, minlength=item_num)[:item_num].to(torch.int32)
    tu, ti = interactions_from_frames(test_data_df)
    test_pos = pos_csr_keep_duplicates(tu, ti, user_num, dev)
    keys = list(lists.keys())
    sums = torch.zeros((max(len(keys), 1), 6), dtype=torch.float64, device=dev)
    for n, (model_name, k) in enumerate(keys):
        rec = lists_to_tensor(lists[(model_name, k)], user_num, int(k), dev)
        ops.topk_metrics(rec, item_num, test_pos, cooc, deg, out=sums[n])
    host = sums.cpu()
    sheets: dict = {}
    for n, (model_name, k) in enumerate(keys):
        m = ops.metrics_from_sums(host[n].tolist(), user_num, int(k))
        sheets.setdefault(int(k), []).append({"Model": model_name, "P": m["precision"], "R": m["recall"], "F1": m["f1"],
                                              "NDCG": m["ndcg"], "H": m["H"], "I": m["I"]})
    frames = {k: pd.DataFrame(rows) for k, rows in sheets.items()}
    if save_path is not None:
        for k, df in frames.items():
            df.to_csv(save_path + f"model_evaluation_results_{k}.csv", index=False)
    return frames
