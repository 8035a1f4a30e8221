"""X3: the reference's main.py call sequence (read CSVs -> recommend* by cfg.MODEL["name"] -> six test metrics)
executed END TO END on the GPU for all seven model names, through tests/_main_sequence.py (a restatement of
main.py:25-106 — /root/reference itself cannot travel to the GPU box).  Every returned list is checked against
the oracle on the same inputs, and the six metrics against an independent NumPy evaluation of the SAME lists
(+-1e-5 after the reference's 5-dp rounding).  BASELINE config 1 (LightGCN, ML-100K shape) is the first case."""
import random

import numpy as np
import pandas as pd
import pytest
import torch

import _main_sequence
import _metrics_numpy as MN
import _stub_const
from _parity import assert_topk_parity
from oracle import lightgcn_oracle as LO
from oracle import spread_oracle as SO

pytestmark = pytest.mark.gpu


def _prepare(cfg, shape):
    from lgcnhs_b200.synth import synth_shape

    d = synth_shape(shape)
    splits = d.split()
    cfg.PREPROCESSING["save_path"] = cfg.PREPROCESSING["save_path"]
    feats = _main_sequence.write_preprocessed(cfg.PREPROCESSING["save_path"], d, splits)
    return d, splits, feats


def _check_metrics(res, k):
    rec = res["recommendations"].numpy()
    acc = MN.accurate_metrics(res["test_pos"], rec, k)
    div = MN.diversity_metrics(rec, res["item_degree_dict"], res["interaction_mat"], k)
    assert np.allclose(res["accurate"], acc, atol=1e-5, equal_nan=True), (res["accurate"], acc)
    assert np.allclose(res["diversity"], div, atol=1e-5, equal_nan=True), (res["diversity"], div)


@pytest.mark.parametrize("name,shape,dataset", [("LightGCN", "ml-100k", "movielens"), ("LightGCNOpti", "small", "movielens")])
def test_main_sequence_lightgcn_family(dev, name, shape, dataset):
    """Config 1: trainLightGCN (a few iterations) -> pickle -> recommendForAllUser -> metrics, all on the device."""
    cfg = _stub_const.install(model=name, dataset=dataset, k=20, epochs=6)
    random.seed(0)
    torch.manual_seed(0)
    d, (tr, va, te), _ = _prepare(cfg, shape)
    res = _main_sequence.run(cfg)
    U, M, k = res["user_num"], res["item_num"], 20
    assert (U, M) == (d.n_users, d.n_items)
    rec = res["recommendations"].numpy()
    assert rec.shape == (U, k) and all(isinstance(v, list) and isinstance(v[0], int) for v in res["rec"].values())
    model = torch.load(cfg.MODEL["save_path"] + f"{k}_{name}.pth", weights_only=False)
    e_tr = torch.from_numpy(np.stack([d.users[tr], d.items[tr]]))
    e_va = torch.from_numpy(np.stack([d.users[va], d.items[va]]))
    score = LO.masked_score(model.users_emb.weight.detach().cpu(), model.items_emb.weight.detach().cpu(), e_tr, e_va)
    rv, ri = LO.topk_items(score, k)                                   # recommend.py:86-114 on the trained weights
    assert_topk_parity(rec, ri.numpy(), score.numpy(), f"{name} main sequence", min_checked=0.9)
    _check_metrics(res, k)
    saved = np.load(cfg.RECOMMEND["save_path"] + f"all_user_recommend_dict_{name}_{k}.npy", allow_pickle=True).item()
    assert saved == res["rec"]


@pytest.mark.parametrize("name,dataset", [("HybridS", "movielens"), ("HeatS", "movielens"), ("ProbS", "douban"),
                                          ("ProbS", "movielens"), ("HeatS", "douban")])
def test_main_sequence_spreading_family(dev, name, dataset):
    cfg = _stub_const.install(model=name, dataset=dataset, k=10, lam=0.4)
    d, (tr, va, te), _ = _prepare(cfg, "small")
    res = _main_sequence.run(cfg)
    U, M, k = res["user_num"], res["item_num"], 10
    tv = np.r_[tr, va]
    A = SO.interaction_matrix(U, M, d.users[tv], d.items[tv])
    Gm = SO.get_spreading_general_mat(A)
    lam, unfiltered = 0.4, False                                        # cfg value: recommend.py:74 overrides the argument
    if name == "ProbS" and dataset == "movielens":                      # recommend.py:89-91, 49-50
        lam, Gm, unfiltered = 0.01, Gm.T, True
    if name == "HeatS" and dataset == "douban":                         # recommend.py:99-101
        lam, Gm = 0.99, Gm.T
    F = SO.get_resource(A, SO.hybrids(A, Gm, lam))
    ref_idx, _ = SO.recommend_fast(F, A, k, unfiltered=unfiltered)
    assert_topk_parity(res["recommendations"].numpy(), ref_idx, F, f"{name}@{dataset} main sequence",
                       seen_mask=None if unfiltered else A > 0, min_checked=0.8)
    _check_metrics(res, k)


@pytest.mark.parametrize("name", ["SpreadLightGCN", "SpreadLightGCNOpti"])
def test_main_sequence_fusion_family(dev, name):
    cfg = _stub_const.install(model=name, dataset="movielens", k=10, lam=0.3, epochs=4)
    random.seed(0)
    torch.manual_seed(0)
    d, (tr, va, te), _ = _prepare(cfg, "tiny")
    res = _main_sequence.run(cfg)                                       # trains the LightGCN(Opti) it needs, pickles it
    U, M, k = res["user_num"], res["item_num"], 10
    base = "LightGCNOpti" if name.endswith("Opti") else "LightGCN"
    model = torch.load(cfg.MODEL["save_path"] + f"{k}_{base}.pth", weights_only=False)
    tv = np.r_[tr, va]
    A = SO.interaction_matrix(U, M, d.users[tv], d.items[tv])
    e_tr = torch.from_numpy(np.stack([d.users[tr], d.items[tr]]))
    e_va = torch.from_numpy(np.stack([d.users[va], d.items[va]]))
    Gs = LO.masked_score(model.users_emb.weight.detach().cpu(), model.items_emb.weight.detach().cpu(), e_tr, e_va).numpy()
    F_new = SO.fused_resource(Gs, SO.get_resource(A, SO.hybrids(A, SO.get_spreading_general_mat(A), 0.3)))
    ref_idx, _ = SO.recommend_fast(F_new, A, k)
    assert_topk_parity(res["recommendations"].numpy(), ref_idx, F_new, f"{name} main sequence", seen_mask=A > 0,
                       min_checked=0.1, tol_mult=4.0)
    _check_metrics(res, k)
