"""A stand-in for the reference's const.py for tests that run where /root/reference does not exist
(the GPU box): same attribute names the drop-in modules read (cfg.DATA_SET, MODEL, RECOMMEND, LOG,
PICTURES), values of DevConfig("movielens", "LightGCN") (reference const.py:111-190)."""
import os
import sys
import tempfile
import types


def install(model: str = "LightGCN", dataset: str = "movielens", k: int = 10, lam: float = 0.3, epochs: int = 6):
    root = tempfile.mkdtemp(prefix="lgc_cfg_")
    paths = {n: os.path.join(root, n) + "/" for n in ("log", "preprocess", "recommend", "model", "evaluation", "pictures")}
    for p in paths.values():
        os.makedirs(p, exist_ok=True)
    cfg = types.SimpleNamespace()
    cfg.DATA_SET = dataset
    cfg.LOG = {"file_path": paths["log"]}
    cfg.PREPROCESSING = {"seed": 42, "save_path": paths["preprocess"]}
    cfg.RECOMMEND = {"k": k, "save_path": paths["recommend"]}
    cfg.EVALUATION = {"save_path": paths["evaluation"]}
    cfg.PICTURES = {"save_path": paths["pictures"]}
    hp = {"seed": 42, "embedding_dim": 64, "layers": 3, "lr": 1e-3, "gamma": 0.95, "epochs": epochs, "epoch_per_eval": 2,
          "epoch_per_lr_decay": 2, "batch_size": 1024, "epsilon": 1e-6, "lambda": lam}
    cfg.MODEL = {"name": model, "HyperParameter": hp, "save_path": paths["model"]}
    mod = types.ModuleType("const")
    mod.cfg = cfg
    sys.modules["const"] = mod
    for name in [m for m in sys.modules if m.split(".")[0] in ("model", "utils", "metrics", "processing")]:
        del sys.modules[name]        # re-import the drop-in against this cfg
    return cfg
