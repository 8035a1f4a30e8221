"""Drop-in for /root/reference/model/SpreadLightGCNOpti/recommend.py."""
import pandas as pd

from const import cfg
from lgcnhs_b200 import fusion
from lgcnhs_b200.recommend_common import topk_dict
from model.SpreadLightGCN.recommend import _save, recommendForAllUser  # noqa: F401
from model.SpreadLightGCNOpti.model import getLightGCNOptiModel, getResourceMat  # noqa: F401


def recommendSpreadLightGCNOpti(user_num: int, item_num: int, rating_df: pd.DataFrame, train_data_df: pd.DataFrame,
                                val_data_df: pd.DataFrame, test_data_df: pd.DataFrame,
                                user_features_df: pd.DataFrame, item_features_df: pd.DataFrame) -> dict:
    """reference recommend.py:56-79, device-resident."""
    k = cfg.RECOMMEND["k"]
    lambda_val = cfg.MODEL["HyperParameter"]["lambda"]
    model = getLightGCNOptiModel(user_num, item_num, rating_df, train_data_df, val_data_df, test_data_df,
                             user_features_df, item_features_df, k)[0]
    out = topk_dict(fusion.fused_recommend(model, user_num, item_num, train_data_df, val_data_df, lambda_val, k))
    _save(out)
    return out
