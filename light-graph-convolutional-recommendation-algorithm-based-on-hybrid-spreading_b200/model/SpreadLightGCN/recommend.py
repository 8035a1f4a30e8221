"""Drop-in for /root/reference/model/SpreadLightGCN/recommend.py."""
import numpy as np
import pandas as pd

from const import cfg
from lgcnhs_b200 import fusion, ops
from lgcnhs_b200.recommend_common import interactions_from_frames, topk_dict, topk_from_host_matrix
from model.SpreadLightGCN.model import getLightGCNModel, getResourceMat  # noqa: F401


def _save(d: dict) -> None:
    np.save(cfg.RECOMMEND["save_path"] + "all_user_recommend_dict_" + cfg.MODEL["name"] + "_" +
            str(cfg.RECOMMEND["k"]) + ".npy", dict(d))


def recommendForAllUser(F_new: np.ndarray, user_num: int, train_data_df: pd.DataFrame,
                        val_data_df: pd.DataFrame, k: int) -> dict:
    """Filtered per-user top-k of a host resource matrix (reference recommend.py:18-52)."""
    u, i = interactions_from_frames(train_data_df, val_data_df)
    excl = ops.ExclusionMask.from_pairs(u, i, user_num, F_new.shape[1])
    out = topk_dict(topk_from_host_matrix(F_new[:user_num], k, excl))
    _save(out)
    return out


def recommendSpreadLightGCN(user_num: int, item_num: int, rating_df: pd.DataFrame, train_data_df: pd.DataFrame,
                            val_data_df: pd.DataFrame, test_data_df: pd.DataFrame) -> dict:
    """getResourceMat + recommendForAllUser of the reference (recommend.py:55-75) without the host round
    trip of the (U, M) matrix: score, spreading, Hadamard and top-k stay on the device."""
    k = cfg.RECOMMEND["k"]
    lambda_val = cfg.MODEL["HyperParameter"]["lambda"]
    model = getLightGCNModel(user_num, item_num, rating_df, train_data_df, val_data_df, test_data_df, k)[0]
    out = topk_dict(fusion.fused_recommend(model, user_num, item_num, train_data_df, val_data_df, lambda_val, k))
    _save(out)
    return out
