"""Drop-in for /root/reference/model/LightGCNOpti/recommend.py."""
import ast

import numpy as np
import pandas as pd
import torch

from const import cfg
from model.LightGCN.recommend import buildGraph, load_or_train, recommendForAllUser  # noqa: F401
from model.LightGCNOpti.train import trainLightGCNOpti


def parse_features(features_df: pd.DataFrame, id_col: str, feat_col: str) -> torch.Tensor:
    """Feature CSV column (python-list literal strings) -> (n, F) float tensor, rows ordered by id
    (reference LightGCNOpti/recommend.py:150-163)."""
    df = features_df.sort_values(by=id_col)
    rows = df[feat_col].apply(lambda r: ast.literal_eval(r) if not isinstance(r, list) else r).tolist()
    return torch.from_numpy(np.array(rows)).float()


def recommendLightGCNOpti(user_num: int, item_num: int, rating_df: pd.DataFrame, train_data_df: pd.DataFrame,
                          val_data_df: pd.DataFrame, test_data_df: pd.DataFrame,
                          user_features_df: pd.DataFrame, item_features_df: pd.DataFrame) -> dict:
    k = cfg.RECOMMEND["k"]
    edge_index, train_edge_index, val_edge_index, test_edge_index = buildGraph(
        user_num, item_num, rating_df, train_data_df, val_data_df, test_data_df)
    user_features = parse_features(user_features_df, "user_id", "user_features")
    item_features = parse_features(item_features_df, "item_id", "item_features")
    model = load_or_train(cfg.MODEL["save_path"] + str(k) + "_LightGCNOpti.pth",
                          lambda: trainLightGCNOpti(user_num, item_num, edge_index, train_edge_index, val_edge_index,
                                                    user_features, item_features),
                          "LightGCNOpti")
    return recommendForAllUser(model, user_num, item_num, train_edge_index, val_edge_index, test_edge_index, k)
