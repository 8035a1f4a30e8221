"""Drop-in for /root/reference/model/SpreadMethod/recommend.py: recommendForAllUser and
recommendSpreadMethod with the reference's signatures, dispatch rules and return convention.

recommendSpreadMethod stays on the device from the interaction list to the top-k ids
(SpreadingEngine: tcgen05 G and F GEMMs, fused scaling, radix-select top-k); the per-user
`np.argsort` + Python `not in` filter loop of reference recommend.py:35-47 (32.8 ms/user) is one
kernel launch."""
import numpy as np
import pandas as pd

from const import cfg
from lgcnhs_b200 import ops
from lgcnhs_b200.recommend_common import engine_from_frames, interactions_from_frames, topk_dict, topk_from_host_matrix


def _save(all_user_recommend_dict: dict) -> None:
    np.save(cfg.RECOMMEND["save_path"] + "all_user_recommend_dict_" + cfg.MODEL["name"] + "_" +
            str(cfg.RECOMMEND["k"]) + ".npy", dict(all_user_recommend_dict))


def _unfiltered() -> bool:
    # movielens + ProbS returns the UNFILTERED head of the ranking (reference recommend.py:48-50)
    return cfg.DATA_SET == "movielens" and cfg.MODEL["name"] == "ProbS"


def recommendForAllUser(F_new: np.ndarray, user_num: int, train_data_df: pd.DataFrame,
                        val_data_df: pd.DataFrame, k: int) -> dict:
    """Per user: rank items by F_new (descending), drop train+val items, keep k (reference recommend.py:18-56)."""
    u, i = interactions_from_frames(train_data_df, val_data_df)
    excl = None if _unfiltered() else ops.ExclusionMask.from_pairs(u, i, user_num, F_new.shape[1])
    out = topk_dict(topk_from_host_matrix(F_new[:user_num], k, excl), as_array_rows=_unfiltered())
    _save(out)
    return out


def recommendSpreadMethod(user_num: int, item_num: int, train_data_df: pd.DataFrame, val_data_df: pd.DataFrame,
                          method: str, lambda_val: float = 0) -> dict:
    """ProbS | HeatS | HybridS — all three go through HybridS(lambda) exactly as in the reference
    (recommend.py:59-115); the `lambda_val` argument is overwritten from cfg (:74) like there."""
    k = cfg.RECOMMEND["k"]
    lambda_val = cfg.MODEL["HyperParameter"]["lambda"]
    if method not in ["ProbS", "HeatS", "HybridS"]:
        raise ValueError(f"Invalid parameter: method={method}，必须为 ProbS | HeatS | HybridS")
    eng = engine_from_frames(user_num, item_num, train_data_df, val_data_df)
    eng.general_w()
    # dataset-specific overrides (recommend.py:89-91, 99-101).  The reference also transposes general_W
    # there; G is symmetric (its transpose differs from it by ulps only), so only lambda changes.
    if method == "ProbS" and cfg.DATA_SET == "movielens":
        lambda_val = 0.01
    elif method == "HeatS" and cfg.DATA_SET == "douban":
        lambda_val = 0.99
    idx, _ = eng.recommend(float(lambda_val), k, filtered=not _unfiltered())
    out = topk_dict(idx, as_array_rows=_unfiltered())
    _save(out)
    return out
