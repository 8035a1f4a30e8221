"""Drop-in for /root/reference/model/SpreadLightGCNOpti/model.py (feature-initialised LightGCN x
hybrid spreading; also imported by findLambda.py:17-18)."""
import numpy as np
import pandas as pd

from const import cfg
from lgcnhs_b200 import fusion
from model.LightGCN.recommend import buildGraph, load_or_train
from model.LightGCNOpti.recommend import parse_features
from model.LightGCNOpti.train import trainLightGCNOpti
from model.SpreadMethod.model import HybridS, getResource, getSpreadingGeneralMat  # noqa: F401
from utils.log import logger
from utils.wrapper import calTimes


def getLightGCNOptiModel(user_num: int, item_num: int, rating_df: pd.DataFrame, train_data_df: pd.DataFrame,
                     val_data_df: pd.DataFrame, test_data_df: pd.DataFrame, user_features_df: pd.DataFrame,
                     item_features_df: pd.DataFrame, k: int) -> tuple:
    """(model, edge_index, train_adj, val_adj, test_adj) — reference model.py:25-94."""
    edge_index, train_edge_index, val_edge_index, test_edge_index = buildGraph(
        user_num, item_num, rating_df, train_data_df, val_data_df, test_data_df)
    user_features = parse_features(user_features_df, "user_id", "user_features")
    item_features = parse_features(item_features_df, "item_id", "item_features")
    model = load_or_train(cfg.MODEL["save_path"] + str(k) + "_LightGCNOpti.pth",
                          lambda: trainLightGCNOpti(user_num, item_num, edge_index, train_edge_index, val_edge_index,
                                                    user_features, item_features),
                          "LightGCNOpti")
    return model, edge_index, train_edge_index, val_edge_index, test_edge_index


getLightGCNModel = getLightGCNOptiModel   # round-1 name of this drop-in, kept as an alias


@calTimes(logger, "分配权重矩阵计算完成")
def getAllocateMat(user_num: int, item_num: int, rating_df: pd.DataFrame, train_data_df: pd.DataFrame,
                   val_data_df: pd.DataFrame, test_data_df: pd.DataFrame, user_features_df: pd.DataFrame,
                   item_features_df: pd.DataFrame, k: int) -> np.ndarray:
    """reference model.py:97-169."""
    model = getLightGCNOptiModel(user_num, item_num, rating_df, train_data_df, val_data_df, test_data_df,
                             user_features_df, item_features_df, k)[0]
    return fusion.allocate_score_device(model, user_num, item_num, train_data_df, val_data_df).cpu().numpy()


@calTimes(logger, "资源扩散矩阵计算完成")
def getHybridSResourceMat(A: np.ndarray, general_W: np.ndarray, lambad_val: float) -> np.ndarray:
    """reference model.py:172-188."""
    return getResource(A, HybridS(A, general_W, lambad_val))


def getResourceMat(user_num: int, item_num: int, rating_df: pd.DataFrame, train_data_df: pd.DataFrame,
                   val_data_df: pd.DataFrame, test_data_df: pd.DataFrame, user_features_df: pd.DataFrame,
                   item_features_df: pd.DataFrame) -> np.ndarray:
    """F_new = G * F (reference model.py:191-243)."""
    k = cfg.RECOMMEND["k"]
    lambda_val = cfg.MODEL["HyperParameter"]["lambda"]
    model = getLightGCNOptiModel(user_num, item_num, rating_df, train_data_df, val_data_df, test_data_df,
                             user_features_df, item_features_df, k)[0]
    return fusion.resource_mat_host(model, user_num, item_num, train_data_df, val_data_df, lambda_val)
