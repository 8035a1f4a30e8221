"""Shared implementation of model/SpreadLightGCN/model.py and model/SpreadLightGCNOpti/model.py:
LightGCN layer-0 score (masked to -1024 at train/val pairs) ⊙ hybrid-spreading resource matrix
(/root/reference/model/SpreadLightGCN/model.py:55-153, SpreadLightGCNOpti/model.py:97-243)."""
from __future__ import annotations

import numpy as np
import pandas as pd
import torch

from . import ops
from .recommend_common import cuda_device, engine_from_frames, interactions_from_frames


def allocate_score_device(model, user_num: int, item_num: int, train_data_df: pd.DataFrame,
                          val_data_df: pd.DataFrame) -> torch.Tensor:
    """G_score = e_u^0 . e_i^0^T with train and val pairs set to -1024, (U, M) fp32 on the device
    (reference getAllocateMat, SpreadLightGCN/model.py:74-102)."""
    dev = cuda_device()
    model = model.to(dev)
    xu = model.users_emb.weight.detach().contiguous()
    xi = model.items_emb.weight.detach().contiguous()
    u, i = interactions_from_frames(train_data_df, val_data_df)
    seen = ops.seen_csr(u, i, user_num, item_num)
    ld = (item_num + 3) // 4 * 4
    out = torch.empty((user_num, ld), dtype=torch.float32, device=dev)[:, :item_num]
    blk = 8192
    for u0 in range(0, user_num, blk):
        u1 = min(u0 + blk, user_num)
        ops.score_block(xu, xi, u0, u1, seen, fill=-float(1 << 10), out=out[u0:u1])
    return out


def fused_recommend(model, user_num: int, item_num: int, train_data_df: pd.DataFrame, val_data_df: pd.DataFrame,
                    lambda_val: float, k: int) -> torch.Tensor:
    """top-k of (G_score ⊙ A.HybridS(lambda)) with train+val items filtered, all on the device.  G_score is never
    materialised: the layer-0 score tiles are multiplied with F and selected inside one kernel (lgc_score_topk)."""
    dev = cuda_device()
    model = model.to(dev)
    xu = model.users_emb.weight.detach().contiguous()
    xi = model.items_emb.weight.detach().contiguous()
    u, i = interactions_from_frames(train_data_df, val_data_df)
    seen = ops.seen_csr(u, i, user_num, item_num)
    eng = ops.SpreadingEngine(user_num, item_num, u, i)
    eng.general_w()
    eng.scale(float(lambda_val))
    F = eng.resource()
    idx, _ = ops.score_topk(xu, xi, k, seen, fill=-float(1 << 10), exclude_seen=True, mul=F, want_values=False)
    return idx


def resource_mat_host(model, user_num: int, item_num: int, train_data_df, val_data_df, lambda_val: float) -> np.ndarray:
    """F_new = G * F as a host float64 array (reference getResourceMat return value)."""
    G_score = allocate_score_device(model, user_num, item_num, train_data_df, val_data_df)
    eng = engine_from_frames(user_num, item_num, train_data_df, val_data_df)
    eng.general_w()
    eng.scale(float(lambda_val))
    F = eng.resource()
    from ._lib import check, lib

    check(lib().hs_hadamard(F.data_ptr(), G_score.data_ptr(), user_num, item_num, int(F.stride(0)),
                            int(G_score.stride(0)), torch.cuda.current_stream().cuda_stream), "hadamard")
    return F.double().cpu().numpy()
