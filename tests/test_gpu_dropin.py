"""GPU parity through the reference-facing module surface (model.*, utils.*), driven by a stand-in
cfg (tests/_stub_const.py) because /root/reference does not exist on the GPU box."""
import os
import random

import numpy as np
import pandas as pd
import pytest
import torch

import _stub_const
from oracle import lightgcn_oracle as LO
from oracle import spread_oracle as SO
from test_gpu_propagation import assert_close

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _frames(name):
    from lgcnhs_b200.synth import synth_shape

    d = synth_shape(name)
    tr, va, te = d.split()
    df = lambda idx: pd.DataFrame({"user_id": d.users[idx], "item_id": d.items[idx], "rating": 5,  # noqa: E731
                                   "rating_time": "2024-01-01"})
    return d, df(np.arange(d.users.size)), df(tr), df(va), df(te), (tr, va, te)


def test_golden_lightgcn_forward_backward(dev):
    """CUDA LightGCN module vs the vectors recorded from the reference's own model.py / loss.py."""
    _stub_const.install()
    from model.LightGCN.loss import BPRLoss
    from model.LightGCN.model import LightGCN

    z = np.load(os.path.join(G, "lightgcn_tiny.npz"))
    m = LightGCN(96, 160, 64, 3)
    with torch.no_grad():
        m.users_emb.weight.copy_(torch.from_numpy(z["users_w"]))
        m.items_emb.weight.copy_(torch.from_numpy(z["items_w"]))
    m = m.to(dev)
    adj = torch.from_numpy(z["adj"]).to(dev)
    uf, u0, itf, i0 = m.forward(adj)
    assert uf.shape == (96, 64) and itf.shape == (160, 64) and u0 is m.users_emb.weight
    assert_close(uf, torch.from_numpy(z["users_final"]), "users_final vs reference")
    assert_close(itf, torch.from_numpy(z["items_final"]), "items_final vs reference")
    u, p, n = (torch.from_numpy(z[k]).to(dev) for k in ("bpr_u", "bpr_p", "bpr_n"))
    loss = BPRLoss(uf[u], u0[u], itf[p], i0[p], itf[n], i0[n], 1e-6)
    assert abs(loss.item() - float(z["bpr_loss"])) <= 1e-5 * abs(float(z["bpr_loss"]))
    loss.backward()
    assert_close(m.users_emb.weight.grad, torch.from_numpy(z["grad_users"]), "dL/d users_emb vs reference")
    assert_close(m.items_emb.weight.grad, torch.from_numpy(z["grad_items"]), "dL/d items_emb vs reference")
    # the module stays picklable the way the reference saves it (torch.save(model))
    import io

    buf = io.BytesIO()
    torch.save(m, buf)
    buf.seek(0)
    m2 = torch.load(buf, weights_only=False)
    assert torch.equal(m2.users_emb.weight, m.users_emb.weight)


def test_golden_spreading_numpy_api(dev):
    """model.SpreadMethod.model functions (NumPy float64 in/out) vs the reference's outputs."""
    _stub_const.install()
    from model.SpreadMethod import model as SM

    z = np.load(os.path.join(G, "spread_tiny.npz"))
    sel = np.r_[z["train"], z["val"]]
    A = SO.interaction_matrix(96, 160, z["users"][sel], z["items"][sel])
    Gm = SM.getSpreadingGeneralMat(A)
    assert Gm.dtype == np.float64 and Gm.shape == (160, 160)
    assert_close(torch.from_numpy(Gm), torch.from_numpy(z["G"]), "getSpreadingGeneralMat vs reference")
    for lam in (0.0, 0.3, 1.0):
        W = SM.HybridS(A, z["G"], lam)
        assert_close(torch.from_numpy(W), torch.from_numpy(z[f"W_{lam}"]), f"HybridS({lam}) vs reference")
        F = SM.getResource(A, z[f"W_{lam}"])
        assert_close(torch.from_numpy(F), torch.from_numpy(z[f"F_{lam}"]), f"getResource({lam}) vs reference")
    assert_close(torch.from_numpy(SM.ProbS(A, z["G"])), torch.from_numpy(z["W_1.0"]), "ProbS == HybridS(1)")
    assert_close(torch.from_numpy(SM.HeatS(A, z["G"])), torch.from_numpy(z["W_0.0"]), "HeatS == HybridS(0)")


@pytest.mark.parametrize("name,lam", [("tiny", 0.3), ("small", 0.3)])
def test_golden_recommend_spread_method(dev, name, lam):
    """recommendSpreadMethod (device resident) vs the reference's recommendForAllUser output."""
    cfg = _stub_const.install(model="HybridS", lam=lam, k={"tiny": 10, "small": 20}[name])
    from model.SpreadMethod.recommend import recommendForAllUser, recommendSpreadMethod

    z = np.load(os.path.join(G, f"spread_{name}.npz"))
    U, M = {"tiny": (96, 160), "small": (300, 500)}[name]
    mk = lambda idx: pd.DataFrame({"user_id": z["users"][idx], "item_id": z["items"][idx]})  # noqa: E731
    train_df, val_df = mk(z["train"]), mk(z["val"])
    k = cfg.RECOMMEND["k"]
    out = recommendSpreadMethod(U, M, train_df, val_df, "HybridS")
    assert sorted(out.keys()) == list(range(U)) and all(len(v) == k for v in out.values())
    ref, F = z[f"rec_{lam}"], z[f"F_{lam}"]
    got = np.array([out[u] for u in range(U)])
    fv = np.take_along_axis(F, ref, axis=1)
    assert np.allclose(np.take_along_axis(F, got, axis=1), fv, rtol=1e-5, atol=1e-7 * np.abs(F).max())   # score at rank
    gap = np.ones_like(ref, dtype=bool)
    tol = 2 * (1e-5 * np.abs(fv) + 1e-7 * np.abs(F).max())
    gap[:, 1:] &= (fv[:, :-1] - fv[:, 1:]) > tol[:, 1:]
    gap[:, :-1] &= (fv[:, :-1] - fv[:, 1:]) > tol[:, :-1]
    gap[:, -1] = False
    assert np.array_equal(got[gap], ref[gap])                    # identical ids except at float ties
    assert gap.mean() > 0.5
    saved = np.load(cfg.RECOMMEND["save_path"] + f"all_user_recommend_dict_HybridS_{k}.npy", allow_pickle=True).item()
    assert saved.keys() == out.keys()
    # host-matrix entry point (reference signature): reference F in, reference lists out
    out2 = recommendForAllUser(F, U, train_df, val_df, k)
    got2 = np.array([out2[u] for u in range(U)])
    assert np.array_equal(got2[gap], ref[gap])


def test_probs_movielens_is_unfiltered(dev):
    cfg = _stub_const.install(model="ProbS", dataset="movielens", lam=1.0, k=10)
    from model.SpreadMethod.recommend import recommendSpreadMethod

    d, rating, train_df, val_df, test_df, _ = _frames("tiny")
    out = recommendSpreadMethod(d.n_users, d.n_items, train_df, val_df, "ProbS")
    seen = set(zip(pd.concat([train_df, val_df]).user_id, pd.concat([train_df, val_df]).item_id))
    assert any((u, int(i)) in seen for u, items in out.items() for i in items)      # quirk kept: seen items appear
    assert isinstance(out[0], np.ndarray)


def test_fused_trainer_matches_autograd_adam(dev):
    """3 optimisation steps with injected triplets vs torch autograd + torch.optim.Adam on the oracle."""
    _stub_const.install()
    from lgcnhs_b200.synth import bipartite_adj
    from lgcnhs_b200.trainer import FusedBPRTrainer
    from model.LightGCN.model import LightGCN

    d, *_ , (tr, va, te) = _frames("small")
    adj = torch.from_numpy(bipartite_adj(d.n_users, d.users[tr], d.items[tr]))
    torch.manual_seed(42)
    m = LightGCN(d.n_users, d.n_items, 64, 3)
    uw = m.users_emb.weight.detach().clone().requires_grad_()
    iw = m.items_emb.weight.detach().clone().requires_grad_()
    opt = torch.optim.Adam([uw, iw], lr=1e-2)
    m = m.to(dev)
    trainer = FusedBPRTrainer(m, adj.to(dev), lr=1e-2, eps_reg=1e-4)
    trainer.debug_keep_grad = True
    assert m.users_emb.weight.data_ptr() == trainer.X0.data_ptr()          # the module's weights are views of one table
    g = torch.Generator().manual_seed(9)
    for step in range(3):
        u = torch.randint(d.n_users, (512,), generator=g)
        p = torch.randint(d.n_items, (512,), generator=g)
        n = torch.randint(d.n_items, (512,), generator=g)
        ref_loss, gu, gi = LO.train_step(uw, iw, adj, 3, (u, p, n), 1e-4, opt)
        loss = trainer.step(u.to(dev), p.to(dev), n.to(dev))
        # step 0 starts from identical weights: tight check.  Later steps start from weights that already
        # differ by Adam-amplified rounding noise (see below), so only the drift bound applies.
        ltol = 1e-5 if step == 0 else 1e-4
        assert abs(loss[0].item() - ref_loss.item()) <= ltol * abs(ref_loss.item()) + 1e-7
        if step == 0:
            assert_close(trainer.grad_total[: d.n_users], gu, "step 0 dL/d users", sum_abs=gu.abs() + 1e-6)
            assert_close(trainer.grad_total[d.n_users:], gi, "step 0 dL/d items", sum_abs=gi.abs() + 1e-6)
            trainer.debug_keep_grad = False
        assert float(trainer.gE.abs().max()) == 0.0 and float(trainer.gX.abs().max()) == 0.0   # cleaned row-wise after the step
        # Adam's step is lr * m/(sqrt(v)+1e-8): for |g| ~ 1e-8 it amplifies fp32 summation-order noise in g
        # (dw/dg ~ lr/eps), so weights are compared to 5e-4 of one step size; Adam itself is checked
        # bit-tight on identical gradients in test_gpu_train_ops.py::test_adam_matches_torch
        for w, r, nm in ((m.users_emb.weight, uw, "users"), (m.items_emb.weight, iw, "items")):
            err = (w.detach().cpu() - r.detach()).abs().max().item()
            assert err <= 5e-4 * 1e-2 * (step + 1), f"step {step} {nm} weights after Adam: max err {err:.3e}"


def test_train_and_recommend_lightgcn_end_to_end(dev):
    """trainLightGCN -> pickle -> recommendLightGCN through the reference-named entry points."""
    cfg = _stub_const.install(model="LightGCN", k=10, epochs=6)
    random.seed(0)
    torch.manual_seed(0)
    from model.LightGCN.recommend import recommendLightGCN

    d, rating, train_df, val_df, test_df, _ = _frames("small")
    out = recommendLightGCN(d.n_users, d.n_items, rating, train_df, val_df, test_df)
    assert sorted(out.keys()) == list(range(d.n_users)) and all(len(v) == 10 and isinstance(v[0], int) for v in out.values())
    seen = set(zip(pd.concat([train_df, val_df]).user_id.tolist(), pd.concat([train_df, val_df]).item_id.tolist()))
    assert not any((u, i) in seen for u, items in out.items() for i in items)
    assert os.path.exists(cfg.MODEL["save_path"] + "10_LightGCN.pth")
    csv = pd.read_csv(cfg.PICTURES["save_path"] + "LightGCN_10_val_metrics.csv")
    assert list(csv.columns) == ["iters", "train_loss", "val_loss", "val_precision", "val_recall", "val_f1", "val_ndcg",
                                 "val_H", "val_I"] and len(csv) == 3 and np.isfinite(csv.to_numpy()).all()
    # second call loads the pickle instead of training and reproduces the lists
    model = torch.load(cfg.MODEL["save_path"] + "10_LightGCN.pth", weights_only=False)
    ref = LO.masked_score(model.users_emb.weight.detach().cpu(), model.items_emb.weight.detach().cpu(),
                          torch.from_numpy(np.stack([train_df.user_id.values, train_df.item_id.values])),
                          torch.from_numpy(np.stack([val_df.user_id.values, val_df.item_id.values])))
    rv, ri = LO.topk_items(ref, 10)
    got = np.array([recommendLightGCN(d.n_users, d.n_items, rating, train_df, val_df, test_df)[u] for u in range(3)])
    assert np.array_equal(got, ri[:3].numpy())


def test_fusion_recommend_matches_oracle(dev):
    """SpreadLightGCN: (layer-0 score masked to -1024) * (A . HybridS) -> filtered top-k."""
    cfg = _stub_const.install(model="SpreadLightGCN", lam=0.3, k=10, epochs=4)
    random.seed(0)
    from model.SpreadLightGCN.model import getAllocateMat, getResourceMat
    from model.SpreadLightGCN.recommend import recommendSpreadLightGCN

    d, rating, train_df, val_df, test_df, _ = _frames("tiny")
    out = recommendSpreadLightGCN(d.n_users, d.n_items, rating, train_df, val_df, test_df)     # trains + pickles
    model = torch.load(cfg.MODEL["save_path"] + "10_LightGCN.pth", weights_only=False)
    both = pd.concat([train_df, val_df])
    A = SO.interaction_matrix(d.n_users, d.n_items, both.user_id, both.item_id)
    e_tr = torch.from_numpy(np.stack([train_df.user_id.values, train_df.item_id.values]))
    e_va = torch.from_numpy(np.stack([val_df.user_id.values, val_df.item_id.values]))
    Gs = LO.masked_score(model.users_emb.weight.detach().cpu(), model.items_emb.weight.detach().cpu(), e_tr, e_va).numpy()
    F = SO.get_resource(A, SO.hybrids(A, SO.get_spreading_general_mat(A), 0.3))
    Fn = SO.fused_resource(Gs, F)
    assert_close(torch.from_numpy(getAllocateMat(d.n_users, d.n_items, rating, train_df, val_df, test_df, 10)),
                 torch.from_numpy(Gs), "getAllocateMat")
    got_F = getResourceMat(d.n_users, d.n_items, rating, train_df, val_df, test_df)
    assert_close(torch.from_numpy(got_F), torch.from_numpy(Fn), "getResourceMat = G * F",
                 sum_abs=torch.from_numpy(np.abs(Gs) * F))
    fi, fv = SO.recommend_fast(Fn, A, 10)
    got = np.array([out[u] for u in range(d.n_users)])
    assert np.allclose(np.take_along_axis(Fn, got, 1), fv, rtol=1e-4, atol=1e-6 * np.abs(Fn).max())


def test_get_embedding_for_bpr_matches_oracle(dev, monkeypatch):
    """P4 getEmbeddingForBPR (reference train.py:26-59): forward + six row gathers, with the mini-batch injected so
    both sides use the same (u, pos, neg); loss and gradients through the six tensors vs autograd on the oracle."""
    _stub_const.install()
    import model.LightGCN.train as T
    from lgcnhs_b200.synth import bipartite_adj
    from model.LightGCN.loss import BPRLoss
    from model.LightGCN.model import LightGCN

    d, *_, (tr, va, te) = _frames("small")
    adj = torch.from_numpy(bipartite_adj(d.n_users, d.users[tr], d.items[tr]))
    g = torch.Generator().manual_seed(3)
    B = 256
    rows = torch.randint(tr.size, (B,), generator=g)
    u = torch.from_numpy(d.users[tr])[rows]
    p = torch.from_numpy(d.items[tr])[rows]
    n = torch.randint(d.n_items, (B,), generator=g)
    seen = {}
    monkeypatch.setattr(T, "sampleMiniBatch", lambda bs, ei: (seen.setdefault("ei", ei), (u.to(dev), p.to(dev), n.to(dev)))[1])
    torch.manual_seed(42)
    m = LightGCN(d.n_users, d.n_items, 64, 3)
    uw = m.users_emb.weight.detach().clone().requires_grad_()
    iw = m.items_emb.weight.detach().clone().requires_grad_()
    m = m.to(dev)
    six = T.getEmbeddingForBPR(m, d.n_users, d.n_items, adj.to(dev), B, dev)
    # the sampler was handed the (2, E) user->item edge list recovered from the adjacency (train.py:48)
    assert torch.equal(seen["ei"].cpu(), LO.convert_adj_to_edge_index(d.n_users, d.n_items, adj))
    ouf, ou0, oif, oi0 = LO.lightgcn_forward(uw, iw, adj, 3)
    ref6 = (ouf[u], ou0[u], oif[p], oi0[p], oif[n], oi0[n])
    assert len(six) == 6
    for got, ref, nm in zip(six, ref6, ("u_f", "u_0", "p_f", "p_0", "n_f", "n_0")):
        assert got.shape == (B, 64)
        assert_close(got, ref.detach(), f"getEmbeddingForBPR {nm}")
    loss = BPRLoss(*six, 1e-4)
    ref_loss = LO.bpr_loss(*ref6, 1e-4)
    assert abs(loss.item() - ref_loss.item()) <= 1e-5 * abs(ref_loss.item()) + 1e-7
    loss.backward()
    ref_loss.backward()
    assert_close(m.users_emb.weight.grad, uw.grad, "P4 dL/d users", sum_abs=uw.grad.abs() + 1e-6)
    assert_close(m.items_emb.weight.grad, iw.grad, "P4 dL/d items", sum_abs=iw.grad.abs() + 1e-6)


def test_cal_val_loss_matches_oracle(dev, monkeypatch):
    """P9 calValLoss (reference evaluation.py:56-86): propagate over the VAL graph, BPR over ALL val edges with the
    negatives injected, rounded to 5 decimals."""
    _stub_const.install()
    import model.LightGCN.evaluation as E
    from lgcnhs_b200.synth import bipartite_adj
    from model.LightGCN.model import LightGCN

    d, *_, (tr, va, te) = _frames("small")
    val_adj = torch.from_numpy(bipartite_adj(d.n_users, d.users[va], d.items[va]))
    r_mat = LO.convert_adj_to_edge_index(d.n_users, d.n_items, val_adj)          # (2, E_val) user -> item
    g = torch.Generator().manual_seed(5)
    neg = torch.randint(d.n_items, (r_mat.shape[1],), generator=g)
    calls = {}

    def fake_sampler(edge_index, contains_neg_self_loops=True, **kw):
        calls["loops"] = contains_neg_self_loops
        assert torch.equal(edge_index.cpu(), r_mat)
        return edge_index[0], edge_index[1], neg.to(edge_index.device)

    monkeypatch.setattr(E, "structured_negative_sampling", fake_sampler)
    torch.manual_seed(42)
    m = LightGCN(d.n_users, d.n_items, 64, 3)
    uw, iw = m.users_emb.weight.detach().clone(), m.items_emb.weight.detach().clone()
    m = m.to(dev)
    for eps in (1e-6, 1e-3):
        got = E.calValLoss(m, d.n_users, d.n_items, val_adj.to(dev), eps)
        uf, u0, itf, i0 = LO.lightgcn_forward(uw, iw, val_adj, 3)
        u, p = r_mat[0], r_mat[1]
        ref = round(LO.bpr_loss(uf[u], u0[u], itf[p], i0[p], itf[neg], i0[neg], eps).item(), 5)
        assert isinstance(got, float) and abs(got - ref) <= 1e-5 + 1e-5 * abs(ref), (got, ref)
    assert calls["loops"] is False                                                # evaluation.py:72
    # and with the real device sampler: finite, negatives valid (never a positive of the user, never the user id)
    monkeypatch.undo()
    from lgcnhs_b200.sampling import check_status, structured_negative_sampling
    uu, pp, nn_ = structured_negative_sampling(r_mat.to(dev), contains_neg_self_loops=False)
    check_status()
    pos = set(zip(r_mat[0].tolist(), r_mat[1].tolist()))
    assert all((a, b) not in pos and a != b for a, b in zip(uu.tolist(), nn_.tolist()))
    assert int(nn_.max()) <= int(r_mat[1].max()) and int(nn_.min()) >= 0          # drawn inside the item-id range
    assert np.isfinite(E.calValLoss(m, d.n_users, d.n_items, val_adj.to(dev), 1e-6))


def test_negative_range_is_items_even_when_users_outnumber_items(dev):
    """with U > M (ML-1M / ML-20M shapes) negatives must be spread over the items, not clamped onto the
    last one; with U < M the draw range equals the reference's max-id + 1."""
    _stub_const.install()
    from lgcnhs_b200.sampling import check_status
    from model.LightGCN.loss import sampleMiniBatch

    g = np.random.default_rng(0)
    U, M, E = 500, 40, 4000
    ei = torch.from_numpy(np.unique(np.stack([g.integers(0, U, E), g.integers(0, M, E)]), axis=1)).to(dev)
    torch.manual_seed(1)
    u, p, n = sampleMiniBatch(20000, ei)
    check_status()
    assert int(n.max()) < M and int(n.min()) >= 0
    cnt = torch.bincount(n, minlength=M).float()
    assert cnt.max() < 3.0 * cnt.mean()                      # no pile-up on one item (was ~81% at the ML-20M shape)
    pos = set(zip(ei[0].tolist(), ei[1].tolist()))
    assert not any((a, b) in pos for a, b in zip(u.tolist(), n.tolist()))


def test_heats_douban_branch_matches_oracle(dev):
    """S5 dispatch quirk: HeatS on the douban dataset runs HybridS with lambda = 0.99 on the TRANSPOSED general
    matrix (reference SpreadMethod/recommend.py:99-101)."""
    cfg = _stub_const.install(model="HeatS", dataset="douban", lam=0.5, k=10)
    from model.SpreadMethod.recommend import recommendSpreadMethod

    d, rating, train_df, val_df, test_df, _ = _frames("tiny")
    out = recommendSpreadMethod(d.n_users, d.n_items, train_df, val_df, "HeatS")
    both = pd.concat([train_df, val_df])
    A = SO.interaction_matrix(d.n_users, d.n_items, both.user_id, both.item_id)
    Gm = SO.get_spreading_general_mat(A)
    F = SO.get_resource(A, SO.hybrids(A, Gm.T, 0.99))                             # recommend.py:100-101
    ref_idx, _ = SO.recommend_fast(F, A, 10)
    from _parity import assert_topk_parity
    got = np.array([out[u] for u in range(d.n_users)])
    assert_topk_parity(got, ref_idx, F, "HeatS@douban", seen_mask=A > 0, min_checked=0.5)
    # and it is NOT what lambda = 0.5 (the cfg value) or plain HeatS (lambda = 0) would give
    F0 = SO.get_resource(A, SO.hybrids(A, Gm, 0.0))
    assert not np.array_equal(SO.recommend_fast(F0, A, 10)[0], got)
