"""The reference's lambda search (/root/reference/findLambda.py:67-129) as one device-resident sweep.

findLambda.py builds A, G_score = getAllocateMat(...) and general_W once and then, for 101 lambda values, runs
HybridS + A.W on the CPU, multiplies with G_score, ranks every user with the argsort + Python filter loop and
evaluates six metrics with O(U^2) Python loops.  Here the same quantities come out of SpreadingEngine.sweep: the two
tensor-core GEMM operands are packed once, every lambda is scale -> A.W -> fused (layer-0 score x F) top-k -> metric
kernels, and the whole (n_lambda, 6) table crosses PCIe once.  The CSV has the reference's columns and file name
(`lambda_evaluation_<k>.csv`, findLambda.py:119-129)."""
from __future__ import annotations

from typing import Optional, Sequence

import numpy as np
import pandas as pd
import torch

from . import ops
from .recommend_common import cuda_device, interactions_from_frames


def find_lambda(user_num: int, item_num: int, train_data_df: pd.DataFrame, val_data_df: pd.DataFrame,
                test_data_df: pd.DataFrame, k: int, model=None, lambdas: Optional[Sequence[float]] = None,
                save_dir: Optional[str] = None, lists_out: Optional[list] = None) -> pd.DataFrame:
    """DataFrame(lambda, precision, recall, f1, ndcg, H, I) over the lambda grid (default np.arange(0, 1.01, 0.01),
    findLambda.py:83).  `model` (a trained LightGCN / LightGCNOpti module) selects the fusion recommender
    SpreadLightGCN(Opti) as in findLambda.py:79-98; None sweeps plain HybridS (the commented-out alternative, :100)."""
    dev = cuda_device()
    lambdas = np.arange(0, 1 + 0.01, 0.01).tolist() if lambdas is None else [float(x) for x in lambdas]
    u, i = interactions_from_frames(train_data_df, val_data_df)
    eng = ops.SpreadingEngine(user_num, item_num, u, i)
    tu, ti = interactions_from_frames(test_data_df)
    test_pos = ops.seen_csr(tu, ti, user_num, item_num)
    layer0 = None
    if model is not None:
        model = model.to(dev)
        layer0 = (model.users_emb.weight.detach().contiguous(), model.items_emb.weight.detach().contiguous(),
                  ops.seen_csr(u, i, user_num, item_num))
    _, res = eng.sweep(lambdas, k, test_pos, filtered=True, layer0=layer0, lists_out=lists_out)
    frame = pd.DataFrame({"lambda": lambdas, **{c: [m[c] for m in res] for c in ("precision", "recall", "f1", "ndcg", "H", "I")}})
    if save_dir is not None:
        frame.to_csv(save_dir + "lambda_evaluation_" + str(k) + ".csv", index=False)
    return frame
