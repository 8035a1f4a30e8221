"""Drop-in for /root/reference/model/SpreadLightGCN/model.py (LightGCN x hybrid spreading fusion)."""
import numpy as np
import pandas as pd

from const import cfg
from lgcnhs_b200 import fusion
from model.LightGCN.recommend import buildGraph, load_or_train
from model.LightGCN.train import trainLightGCN
from model.SpreadMethod.model import HybridS, getResource, getSpreadingGeneralMat  # noqa: F401
from utils.log import logger
from utils.wrapper import calTimes


def getLightGCNModel(user_num: int, item_num: int, rating_df: pd.DataFrame, train_data_df: pd.DataFrame,
                     val_data_df: pd.DataFrame, test_data_df: pd.DataFrame, k: int) -> tuple:
    """(model, edge_index, train_adj, val_adj, test_adj) — load the pickled LightGCN or train it
    (reference model.py:22-53)."""
    edge_index, train_edge_index, val_edge_index, test_edge_index = buildGraph(
        user_num, item_num, rating_df, train_data_df, val_data_df, test_data_df)
    model = load_or_train(cfg.MODEL["save_path"] + str(k) + "_LightGCN.pth",
                          lambda: trainLightGCN(user_num, item_num, edge_index, train_edge_index, val_edge_index),
                          "LightGCN")
    return model, edge_index, train_edge_index, val_edge_index, test_edge_index


@calTimes(logger, "分配权重矩阵计算完成")
def getAllocateMat(user_num: int, item_num: int, rating_df: pd.DataFrame, train_data_df: pd.DataFrame,
                   val_data_df: pd.DataFrame, test_data_df: pd.DataFrame, k: int) -> np.ndarray:
    """(U, M) fp32 layer-0 score with train/val pairs at -1024 (reference model.py:55-104)."""
    model = getLightGCNModel(user_num, item_num, rating_df, train_data_df, val_data_df, test_data_df, k)[0]
    return fusion.allocate_score_device(model, user_num, item_num, train_data_df, val_data_df).cpu().numpy()


@calTimes(logger, "资源扩散矩阵计算完成")
def getHybridSResourceMat(A: np.ndarray, general_W: np.ndarray, lambad_val: float) -> np.ndarray:
    """F = A . HybridS(A, general_W, lambda) (reference model.py:106-120)."""
    return getResource(A, HybridS(A, general_W, lambad_val))


def getResourceMat(user_num: int, item_num: int, rating_df: pd.DataFrame, train_data_df: pd.DataFrame,
                   val_data_df: pd.DataFrame, test_data_df: pd.DataFrame) -> np.ndarray:
    """F_new = G * F (reference model.py:122-153), computed on the device and returned as float64."""
    k = cfg.RECOMMEND["k"]
    lambda_val = cfg.MODEL["HyperParameter"]["lambda"]
    model = getLightGCNModel(user_num, item_num, rating_df, train_data_df, val_data_df, test_data_df, k)[0]
    return fusion.resource_mat_host(model, user_num, item_num, train_data_df, val_data_df, lambda_val)
