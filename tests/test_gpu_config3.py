"""GPU parity for BASELINE config 3: LightGCNOpti (feature-initialised e^0) and SpreadLightGCNOpti (layer-0 score
(x) hybrid spreading), through the reference-named modules.

* tiny shape: against tests/golden/opti_tiny.npz, recorded from the reference's OWN
  model/LightGCNOpti/model.py + model/SpreadLightGCNOpti/{model,recommend}.py run in the build container;
* Douban shape (640 x 16 000, 64 k interactions — SURVEY.md §8's assumed post-filter size): against the oracle at
  full size (NumPy float64 like the reference: G is 16 000^2).
"""
import os

import numpy as np
import pandas as pd
import pytest
import torch

import _stub_const
from _parity import assert_close_np, assert_topk_parity
from oracle import lightgcn_oracle as LO
from oracle import spread_oracle as SO
from test_cpu_opti import _feature_frames
from test_gpu_propagation import assert_close

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _df(users, items, idx):
    return pd.DataFrame({"user_id": users[idx], "item_id": items[idx], "rating": 5, "rating_time": "2024-01-01 00:00:00"})


def test_lightgcn_opti_forward_matches_reference_golden(dev):
    _stub_const.install(model="LightGCNOpti")
    from model.LightGCNOpti.model import LightGCNOpti

    z = np.load(os.path.join(G, "opti_tiny.npz"))
    torch.manual_seed(42)
    m = LightGCNOpti(96, 160, 64, 3, torch.from_numpy(z["user_feat"]).float(), torch.from_numpy(z["item_feat"]).float())
    m = m.to(dev)
    uf, u0, itf, i0 = m.forward(torch.from_numpy(z["adj"]).to(dev))
    assert u0 is m.users_emb.weight and i0 is m.items_emb.weight
    assert_close(uf, torch.from_numpy(z["users_final"]), "LightGCNOpti users_final vs reference")
    assert_close(itf, torch.from_numpy(z["items_final"]), "LightGCNOpti items_final vs reference")
    # gradients reach the feature-initialised embedding parameters through the fused propagation
    (uf.sum() + itf.sum()).backward()
    assert m.users_emb.weight.grad is not None and torch.isfinite(m.users_emb.weight.grad).all()
    ones = torch.ones(96 + 160, 64)
    ref_g = LO.propagate_layers(ones, torch.from_numpy(z["adj"]), 3)
    ref_g = sum(ref_g) / 4                      # d(sum E)/dX0 = mean_l (A^T)^l 1, A symmetric
    assert_close(torch.cat([m.users_emb.weight.grad, m.items_emb.weight.grad]), ref_g, "dL/dX0 through LightGCNOpti")


def test_spread_lightgcn_opti_pipeline_matches_reference_golden(dev):
    """getAllocateMat / getResourceMat / recommendSpreadLightGCNOpti vs the reference's outputs (tiny)."""
    z = np.load(os.path.join(G, "opti_tiny.npz"))
    k, lam = int(z["k"]), float(z["lam"])
    cfg = _stub_const.install(model="SpreadLightGCNOpti", lam=lam, k=k)
    from model.LightGCNOpti.model import LightGCNOpti
    from model.SpreadLightGCNOpti.model import getAllocateMat, getLightGCNOptiModel, getResourceMat
    from model.SpreadLightGCNOpti.recommend import recommendForAllUser, recommendSpreadLightGCNOpti

    U, M = 96, 160
    users, items = z["users"], z["items"]
    rating, train_df, val_df, test_df = (_df(users, items, np.arange(users.size)), _df(users, items, z["train"]),
                                         _df(users, items, z["val"]), _df(users, items, z["test"]))
    ufd, ifd = _feature_frames(z)
    torch.manual_seed(42)
    m = LightGCNOpti(U, M, 64, 3, torch.from_numpy(z["user_feat"]).float(), torch.from_numpy(z["item_feat"]).float())
    torch.save(m, cfg.MODEL["save_path"] + f"{k}_LightGCNOpti.pth")            # what the reference loads (model.py:79)
    model, ei, tr_adj, va_adj, te_adj = getLightGCNOptiModel(U, M, rating, train_df, val_df, test_df, ufd, ifd, k)
    assert np.array_equal(model.users_emb.weight.detach().cpu().numpy(), z["users_w"])
    assert np.array_equal(tr_adj.cpu().numpy(), z["adj"])
    Gs = getAllocateMat(U, M, rating, train_df, val_df, test_df, ufd, ifd, k)
    assert Gs.dtype == np.float32 and Gs.shape == (U, M)
    assert_close_np(Gs, z["G_score"], "getAllocateMat vs reference")
    assert np.array_equal(Gs == -1024.0, z["G_score"] == -1024.0)               # the masked pairs, exactly
    F_new = getResourceMat(U, M, rating, train_df, val_df, test_df, ufd, ifd)
    assert F_new.dtype == np.float64
    both = pd.concat([train_df, val_df])
    A = SO.interaction_matrix(U, M, both.user_id, both.item_id)
    F = SO.get_resource(A, SO.hybrids(A, SO.get_spreading_general_mat(A), lam))
    # |G*F| error budget: both factors carry 1e-5-relative error -> relative to |G||F| per entry
    assert_close_np(F_new, z["F_new"], "getResourceMat vs reference", extra=2e-5 * np.abs(z["G_score"]) * F)
    out = recommendSpreadLightGCNOpti(U, M, rating, train_df, val_df, test_df, ufd, ifd)
    assert sorted(out.keys()) == list(range(U)) and all(len(v) == k for v in out.values())
    got = np.array([out[u] for u in range(U)])
    assert_topk_parity(got, z["rec"], z["F_new"], "recommendSpreadLightGCNOpti vs reference", seen_mask=A > 0,
                       min_checked=0.3)
    saved = np.load(cfg.RECOMMEND["save_path"] + f"all_user_recommend_dict_SpreadLightGCNOpti_{k}.npy",
                    allow_pickle=True).item()
    assert saved.keys() == out.keys()
    # host-matrix entry point with the reference's own F_new
    out2 = recommendForAllUser(z["F_new"], U, train_df, val_df, k)
    assert_topk_parity(np.array([out2[u] for u in range(U)]), z["rec"], z["F_new"], "recommendForAllUser(F_new)",
                       seen_mask=A > 0, min_checked=0.3)


def test_spread_lightgcn_opti_douban_shape_vs_oracle(dev):
    """Config 3 at the Douban shape: fused (layer-0 score masked to -1024) * (A . HybridS(lambda)) -> filtered top-k,
    against the float64 oracle at full size."""
    cfg = _stub_const.install(model="SpreadLightGCNOpti", dataset="douban", lam=0.3, k=20)
    from lgcnhs_b200 import fusion
    from lgcnhs_b200.synth import synth_shape
    from model.LightGCNOpti.model import LightGCNOpti

    d = synth_shape("douban")
    U, M, k, lam = d.n_users, d.n_items, 20, 0.3
    tr, va, te = d.split()
    frng = np.random.default_rng(3)
    uf = torch.from_numpy(frng.random((U, 29)).astype(np.float32))
    itf = torch.from_numpy(frng.random((M, 31)).astype(np.float32))
    torch.manual_seed(42)
    model = LightGCNOpti(U, M, 64, 3, uf, itf)
    train_df, val_df = _df(d.users, d.items, tr), _df(d.users, d.items, va)
    idx = fusion.fused_recommend(model, U, M, train_df, val_df, lam, k).cpu().numpy()
    tv = np.r_[tr, va]
    A = SO.interaction_matrix(U, M, d.users[tv], d.items[tv])
    Gm = SO.get_spreading_general_mat(A)
    F = SO.get_resource(A, SO.hybrids(A, Gm, lam))
    e_tv = torch.from_numpy(np.stack([d.users[tv], d.items[tv]]))
    Gs = LO.masked_score(model.users_emb.weight.detach().cpu(), model.items_emb.weight.detach().cpu(), e_tv).numpy()
    F_new = SO.fused_resource(Gs, F)
    ref_idx, _ = SO.recommend_fast(F_new, A, k)
    frac = assert_topk_parity(idx, ref_idx, F_new, "douban-shape fused recommend vs oracle", seen_mask=A > 0,
                              min_checked=0.2, tol_mult=4.0)
    assert frac > 0.2
    # the matrices themselves, through the NumPy entry points
    from model.SpreadMethod import model as SM
    Gd = SM.getSpreadingGeneralMat(A)
    assert_close_np(Gd, Gm, "G at the Douban shape vs oracle")
    Fd = fusion.resource_mat_host(model, U, M, train_df, val_df, lam)
    assert_close_np(Fd, F_new, "F_new at the Douban shape vs oracle", extra=2e-5 * np.abs(Gs) * F)
