"""Drop-in for /root/reference/model/LightGCN/recommend.py: buildGraph, recommendForAllUser,
recommendLightGCN with the reference's signatures and return convention (dict{uid: [int]*k})."""
import numpy as np
import pandas as pd
import torch

from const import cfg
from model.LightGCN.evaluation import _topk_layer0
from model.LightGCN.model import LightGCN
from model.LightGCN.train import trainLightGCN
from utils.graph import convertAdjMatrixToEdgeIndex, convertEdgeIndexToAdjMatrix
from utils.log import logger
from utils.wrapper import calTimes


def _edges(df: pd.DataFrame) -> torch.Tensor:
    return torch.stack([torch.tensor(df["user_id"].values, dtype=torch.long),
                        torch.tensor(df["item_id"].values, dtype=torch.long)], dim=0)


@calTimes(logger, "LightGCN图建立完成")
def buildGraph(user_num: int, item_num: int, rating_df: pd.DataFrame, train_data_df: pd.DataFrame,
               val_data_df: pd.DataFrame, test_data_df: pd.DataFrame) -> tuple:
    """(edge_index, train_adj, val_adj, test_adj) — reference recommend.py:22-66."""
    edge_index = _edges(rating_df)
    train_edge_index = convertEdgeIndexToAdjMatrix(user_num, item_num, _edges(train_data_df))
    val_edge_index = convertEdgeIndexToAdjMatrix(user_num, item_num, _edges(val_data_df))
    test_edge_index = convertEdgeIndexToAdjMatrix(user_num, item_num, _edges(test_data_df))
    return edge_index, train_edge_index, val_edge_index, test_edge_index


def _save(all_user_recommend_dict: dict) -> None:
    np.save(cfg.RECOMMEND["save_path"] + "all_user_recommend_dict_" + cfg.MODEL["name"] + "_" +
            str(cfg.RECOMMEND["k"]) + ".npy", all_user_recommend_dict)


def recommendForAllUser(model: LightGCN, user_num: int, item_num: int, train_edge_index: torch.Tensor,
                        val_edge_index: torch.Tensor, test_edge_index: torch.Tensor, k: int) -> dict:
    """top-k of e_u^0 . e_i^0^T with train AND val pairs masked to -1024 (reference recommend.py:68-125)."""
    dev = model.users_emb.weight.device          # buildGraph hands over host tensors: convert on the device
    train_ei = convertAdjMatrixToEdgeIndex(user_num, item_num, train_edge_index.to(dev))
    val_ei = convertAdjMatrixToEdgeIndex(user_num, item_num, val_edge_index.to(dev))
    rec = _topk_layer0(model, user_num, item_num, [train_ei, val_ei], k).cpu().tolist()
    all_user_recommend_dict = {uid: items for uid, items in enumerate(rec)}
    _save(all_user_recommend_dict)
    return all_user_recommend_dict


def load_or_train(path: str, train_fn, what: str):
    """torch.load the pickled module, else train.  The reference's bare `except:` (recommend.py:152)
    also swallowed kernel/import errors and silently retrained; only a missing or unreadable
    checkpoint triggers training here."""
    logger.info(f"正在加载{what}模型")
    try:
        model = torch.load(path, weights_only=False)   # whole-module pickle, torch >= 2.6 needs the flag
        logger.info("模型加载完毕")
    except (FileNotFoundError, EOFError, AttributeError, ModuleNotFoundError) as e:
        logger.info(f"模型加载失败（{type(e).__name__}），正在重新训练模型")
        model = train_fn()
    if torch.cuda.is_available():
        model = model.to(torch.device("cuda", torch.cuda.current_device()))
    return model


def recommendLightGCN(user_num: int, item_num: int, rating_df: pd.DataFrame, train_data_df: pd.DataFrame,
                      val_data_df: pd.DataFrame, test_data_df: pd.DataFrame) -> dict:
    k = cfg.RECOMMEND["k"]
    edge_index, train_edge_index, val_edge_index, test_edge_index = buildGraph(
        user_num, item_num, rating_df, train_data_df, val_data_df, test_data_df)
    model = load_or_train(cfg.MODEL["save_path"] + str(k) + "_LightGCN.pth",
                          lambda: trainLightGCN(user_num, item_num, edge_index, train_edge_index, val_edge_index),
                          "LightGCN")
    return recommendForAllUser(model, user_num, item_num, train_edge_index, val_edge_index, test_edge_index, k)
