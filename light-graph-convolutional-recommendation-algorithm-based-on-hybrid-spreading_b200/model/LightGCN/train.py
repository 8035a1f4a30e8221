"""Drop-in for /root/reference/model/LightGCN/train.py: trainLightGCN / getEmbeddingForBPR with the
reference's signatures, hyper-parameters (cfg), logging, model pickle and metrics CSV — the
step itself runs on the fused CUDA path (lgcnhs_b200.trainer.FusedBPRTrainer)."""
import pandas as pd
import torch

from const import cfg
from lgcnhs_b200.sampling import check_status
from lgcnhs_b200.trainer import FusedBPRTrainer, choose_device, safe_diversity
from metrics.accurate import getAccurateMetrics
from metrics.diversity import getDiversityMetrics
from model.LightGCN.evaluation import calValLoss, getValRecommendations
from model.LightGCN.loss import BPRLoss, sampleMiniBatch  # noqa: F401  (re-exported like the reference)
from model.LightGCN.model import LightGCN
from utils.graph import convertAdjMatrixToEdgeIndex
from utils.log import logger
from utils.picture import plotMetric
from utils.trans import getInteractionMatrixByEdgeIndex, getItemDegreeByUserPosItemDict, getUserItemsDictByEdgeIndex
from utils.wrapper import calTimes

_DENSE_LIMIT = 2 * 10 ** 8   # largest U*M for which the dense float64 interaction matrix is built


def getEmbeddingForBPR(model: LightGCN, user_num: int, item_num: int,
                       train_edge_index: torch.Tensor, batch_size: int, device: torch.device) -> tuple:
    """The six gathered (batch, dim) tensors of reference train.py:26-59, differentiable."""
    users_f, users_0, items_f, items_0 = model.forward(train_edge_index)
    edge_index_to_use = convertAdjMatrixToEdgeIndex(user_num, item_num, train_edge_index)
    u, p, n = sampleMiniBatch(batch_size, edge_index_to_use)
    u, p, n = u.to(device), p.to(device), n.to(device)
    return users_f[u], users_0[u], items_f[p], items_0[p], items_f[n], items_0[n]


def _run_training(model, name: str, user_num: int, item_num: int, train_edge_index, val_edge_index):
    hp = cfg.MODEL["HyperParameter"]
    lr, gamma, epochs = hp["lr"], hp["gamma"], hp["epochs"]
    epoch_per_eval, epoch_per_lr_decay = hp["epoch_per_eval"], hp["epoch_per_lr_decay"]
    batch_size, epsilon_val = hp["batch_size"], hp["epsilon"]
    k = cfg.RECOMMEND["k"]

    model.train()
    trainer = FusedBPRTrainer(model, train_edge_index, lr, epsilon_val)
    hist = {n: [] for n in ("train_loss", "val_loss", "val_precision", "val_recall", "val_f1", "val_ndcg", "val_H", "val_I")}

    # graph-derived constants, computed ONCE (the reference re-derives the first one every step, train.py:48)
    train_ei = convertAdjMatrixToEdgeIndex(user_num, item_num, train_edge_index)
    val_ei = convertAdjMatrixToEdgeIndex(user_num, item_num, val_edge_index)
    train_user_pos_items_dict = getUserItemsDictByEdgeIndex(train_ei)
    val_user_pos_items_dict = getUserItemsDictByEdgeIndex(val_ei)
    train_item_degree_dict = getItemDegreeByUserPosItemDict(train_user_pos_items_dict)
    train_interaction_mat = (getInteractionMatrixByEdgeIndex(user_num, item_num, train_ei)
                             if user_num * item_num <= _DENSE_LIMIT else None)

    for epoch in range(epochs):
        u, p, n = sampleMiniBatch(batch_size, train_ei)
        loss = trainer.step(u.contiguous(), p.contiguous(), n.contiguous())
        if epoch % epoch_per_eval == 0:
            model.eval()
            with torch.no_grad():
                train_loss = round(loss[0].item(), 5)
                check_status()    # sampler range flags of the steps since the last evaluation (no extra sync: .item() above)
                val_loss = calValLoss(model, user_num, item_num, val_edge_index, epsilon_val)
                recommendations = getValRecommendations(model, user_num, item_num, train_edge_index, val_edge_index, k)
                val_precision, val_recall, val_f1, val_ndcg = getAccurateMetrics(val_user_pos_items_dict, recommendations, k)
                val_H, val_I = safe_diversity(recommendations, train_item_degree_dict, train_interaction_mat, k,
                                              getDiversityMetrics)
            for key, v in zip(hist, (train_loss, val_loss, val_precision, val_recall, val_f1, val_ndcg, val_H, val_I)):
                hist[key].append(v)
            logger.info(f"[Iteration {epoch}/{epochs}]" +
                        f"train_loss: {train_loss}, val_loss: {val_loss}, val_precision@{k}: {val_precision}, " +
                        f"val_recall@{k}: {val_recall}, val_f1@{k}: {val_f1}, val_NDCG@{k}: {val_ndcg}, " +
                        f"val_H@{k}: {val_H}, val_I@{k}: {val_I}")
            model.train()
        if epoch % epoch_per_lr_decay == 0 and epoch != 0:
            trainer.decay_lr(gamma)

    torch.save(model, cfg.MODEL["save_path"] + str(k) + "_" + name + ".pth")
    iters = [e * epoch_per_eval for e in range(len(hist["train_loss"]))]
    save_path = cfg.PICTURES["save_path"] + name + "_" + str(k)
    pd.DataFrame({"iters": iters, **hist}).to_csv(save_path + "_val_metrics.csv", index=False)
    for key, ylabel, fname in (("val_precision", "precision", "_precision.png"), ("val_recall", "recall", "_recall.png"),
                               ("val_f1", "F1-score", "_F1-score.png"), ("val_ndcg", "NDCG", "_NDCG.png"),
                               ("val_H", "H", "_H.png"), ("val_I", "I", "_I.png")):
        plotMetric(iters, hist[key], "iteration", ylabel, ylabel + " curves", save_path + fname)
    return model


@calTimes(logger, "模型训练完成")
def trainLightGCN(user_num: int, item_num: int, edge_index: torch.Tensor,
                  train_edge_index: torch.Tensor, val_edge_index: torch.Tensor) -> LightGCN:
    """Train LightGCN with BPR on the train adjacency (reference train.py:62-223)."""
    hp = cfg.MODEL["HyperParameter"]
    device = choose_device()
    logger.info(f"使用设备：{device}")
    torch.manual_seed(hp["seed"])
    model = LightGCN(user_num, item_num, hp["embedding_dim"], hp["layers"]).to(device)
    return _run_training(model, "LightGCN", user_num, item_num, train_edge_index.to(device), val_edge_index.to(device))
