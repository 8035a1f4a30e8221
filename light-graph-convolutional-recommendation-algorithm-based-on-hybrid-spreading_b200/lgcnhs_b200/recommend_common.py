"""Shared host logic of the spreading-family recommenders (SpreadMethod, SpreadLightGCN(Opti))."""
from __future__ import annotations

from collections import defaultdict
from typing import Optional

import numpy as np
import pandas as pd
import torch

from . import ops


def cuda_device() -> torch.device:
    if not torch.cuda.is_available():
        raise RuntimeError("no CUDA device - the B200 drop-in has no CPU fallback")
    return torch.device("cuda", torch.cuda.current_device())


def interactions_from_frames(*frames: pd.DataFrame):
    """(users, items) int64 device tensors of pd.concat(frames) — what the reference turns into the
    dense float64 A with an iterrows loop (utils/trans.py:13-29 via SpreadMethod/recommend.py:81)."""
    dev = cuda_device()
    u = np.concatenate([f["user_id"].to_numpy(dtype=np.int64) for f in frames])
    i = np.concatenate([f["item_id"].to_numpy(dtype=np.int64) for f in frames])
    return torch.from_numpy(u).to(dev), torch.from_numpy(i).to(dev)


def engine_from_frames(user_num: int, item_num: int, *frames: pd.DataFrame) -> ops.SpreadingEngine:
    u, i = interactions_from_frames(*frames)
    return ops.SpreadingEngine(user_num, item_num, u, i)


def topk_dict(idx: torch.Tensor, as_array_rows: bool = False) -> dict:
    """(U, k) device ids -> the reference's defaultdict{uid: [np.int64]*k} (or ndarray rows for the
    unfiltered movielens+ProbS case, SpreadMethod/recommend.py:49-50)."""
    rec = idx.cpu().numpy()
    out = defaultdict(list)
    for uid in range(rec.shape[0]):
        out[uid] = rec[uid] if as_array_rows else list(rec[uid])
    return out


def topk_from_host_matrix(F_new: np.ndarray, k: int, excl: Optional[ops.ExclusionMask]) -> torch.Tensor:
    """Row-wise (filtered) top-k of a HOST matrix: uploaded in user blocks as fp32."""
    dev = cuda_device()
    U, M = F_new.shape
    out = torch.empty((U, k), dtype=torch.int64, device=dev)
    blk = max(1, min(U, (1 << 28) // max(M, 1)))
    for u0 in range(0, U, blk):
        u1 = min(u0 + blk, U)
        S = torch.from_numpy(np.ascontiguousarray(F_new[u0:u1], dtype=np.float32)).to(dev)
        idx, _ = ops.topk_rows(S, k, excl, row_offset=u0, want_values=False)
        out[u0:u1] = idx
    return out
