"""Drop-in for /root/reference/model/LightGCNOpti/train.py: trainLightGCNOpti = trainLightGCN with
the feature-initialised model (reference LightGCNOpti/train.py:62-228)."""
import torch

from const import cfg
from lgcnhs_b200.trainer import choose_device
from model.LightGCN.train import _run_training, getEmbeddingForBPR  # noqa: F401
from model.LightGCNOpti.model import LightGCNOpti
from utils.log import logger
from utils.wrapper import calTimes


@calTimes(logger, "模型训练完成")
def trainLightGCNOpti(user_num: int, item_num: int, edge_index: torch.Tensor, train_edge_index: torch.Tensor,
                      val_edge_index: torch.Tensor, user_features: torch.Tensor,
                      item_features: torch.Tensor) -> LightGCNOpti:
    hp = cfg.MODEL["HyperParameter"]
    device = choose_device()
    logger.info(f"使用设备：{device}")
    torch.manual_seed(hp["seed"])
    model = LightGCNOpti(user_num, item_num, hp["embedding_dim"], hp["layers"], user_features, item_features).to(device)
    return _run_training(model, "LightGCNOpti", user_num, item_num, train_edge_index.to(device),
                         val_edge_index.to(device))
