"""GPU parity: (N1) six metrics of the top-k lists on the device, and the (N4) lambda-sweep driver.

Pinned against the REAL reference: tests/golden/metrics_small.npz holds the outputs of the reference's own
metrics/accurate.py + metrics/diversity.py (oracle/make_golden.py).  Tolerance: the reference rounds every
metric to 5 decimals, so values must agree to 1e-5."""
import os

import numpy as np
import pytest
import torch

from oracle import spread_oracle as S

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _engine(dev, n_users, n_items, users, items):
    from lgcnhs_b200 import ops

    return ops.SpreadingEngine(n_users, n_items, torch.from_numpy(users).to(dev), torch.from_numpy(items).to(dev))


def test_metrics_match_reference_golden(dev):
    from lgcnhs_b200 import ops

    z = np.load(os.path.join(G, "metrics_small.npz"))
    users, items = z["users"], z["items"]
    te, tv = z["test"], np.r_[z["train"], z["val"]]
    U, M, k = 300, 500, 10
    eng = _engine(dev, U, M, users[tv], items[tv])
    C = eng.cooccurrence()
    A = S.interaction_matrix(U, M, users[tv], items[tv])
    assert np.array_equal(C.cpu().numpy().astype(np.int64), (A.T @ A).astype(np.int64))     # exact integers
    pos = ops.seen_csr(torch.from_numpy(users[te]).to(dev), torch.from_numpy(items[te]).to(dev), U, M)
    rec = torch.from_numpy(z["rec"]).to(dev).long().contiguous()
    sums = ops.topk_metrics(rec, M, pos, C, eng.ki).cpu()
    m = ops.metrics_from_sums(sums.tolist(), U, k)
    acc, div = z["accurate"], z["diversity"]
    assert np.allclose([m["precision"], m["recall"], m["f1"], m["ndcg"]], acc, atol=1e-5)
    assert np.allclose([m["H"], m["I"]], div, atol=1e-5)


def test_metrics_duplicates_and_padding(dev):
    """Lists with a repeated item and -1 padding: the reference intersects SETS for H and skips equal ids for I."""
    import _stub_const

    _stub_const.install()
    import _metrics_numpy as MN
    from lgcnhs_b200 import ops
    U, M, k = 40, 60, 6
    g = np.random.default_rng(5)
    users = g.integers(0, U, 600)
    items = g.integers(0, M, 600)
    eng = _engine(dev, U, M, users, items)
    rec = torch.from_numpy(np.stack([g.permutation(M)[:k] for _ in range(U)]))
    rec[3, 4] = rec[3, 1]           # duplicate inside a list
    A = S.interaction_matrix(U, M, users, items)
    deg = {i: int(c) for i, c in enumerate(A.sum(0)) if c > 0}
    H_ref = MN.hamming_distance(rec.numpy(), k)
    I_ref = MN.internal_similarity(rec.numpy(), deg, A, k)
    sums = ops.topk_metrics(rec.to(dev).contiguous(), M, None, eng.cooccurrence(), eng.ki).cpu()
    m = ops.metrics_from_sums(sums.tolist(), U, k)
    assert abs(m["H"] - H_ref) <= 1e-5 and abs(m["I"] - I_ref) <= 1e-5


def test_lambda_sweep_matches_per_lambda_oracle(dev):
    """N4: findLambda.py:83-116 pattern — per lambda HybridS, A.W, filtered top-k, six metrics — against the oracle."""
    import _stub_const

    _stub_const.install()
    import _metrics_numpy as MN
    from _parity import assert_topk_parity
    from lgcnhs_b200 import ops
    from lgcnhs_b200.synth import synth_shape

    d = synth_shape("small")
    tr, va, te = d.split()
    tv = np.r_[tr, va]
    U, M, k = d.n_users, d.n_items, 10
    eng = _engine(dev, U, M, d.users[tv], d.items[tv])
    pos = ops.seen_csr(torch.from_numpy(d.users[te]).to(dev), torch.from_numpy(d.items[te]).to(dev), U, M)
    lams = [0.0, 0.3, 0.85, 1.0]
    lists = []
    _, res = eng.sweep(lams, k, pos, lists_out=lists)
    A = S.interaction_matrix(U, M, d.users[tv], d.items[tv])
    Gm = S.get_spreading_general_mat(A)
    test_dict = {}
    for u, i in zip(d.users[te].tolist(), d.items[te].tolist()):
        test_dict.setdefault(u, []).append(i)
    deg = {i: int(c) for i, c in enumerate(A.sum(0)) if c > 0}
    for lam, m, dev_idx in zip(lams, res, lists):
        F = S.get_resource(A, S.hybrids(A, Gm, lam))
        idx, _ = S.recommend_fast(F, A, k)
        # (1) the lists: identical to the oracle's except at float ties (tie-aware, score at rank always checked)
        got_idx = dev_idx.cpu().numpy()
        assert_topk_parity(got_idx, idx, F, f"sweep lists lambda={lam}", seen_mask=A > 0, min_checked=0.8)
        # (2) the six metrics of exactly those lists, evaluated independently in NumPy (the reference's formulas,
        #     pinned to the reference's outputs in test_cpu_host_logic.py): +-1e-5 after the 5-dp rounding
        p, r, f1, n = MN.accurate_metrics(test_dict, got_idx, k)
        H, I = MN.diversity_metrics(got_idx, deg, A, k)
        got = [m["precision"], m["recall"], m["f1"], m["ndcg"], m["H"], m["I"]]
        assert np.allclose(got, [p, r, f1, n, H, I], atol=1e-5), (lam, got, [p, r, f1, n, H, I])
        # (3) and the oracle's own lists give the same metrics wherever no tie was involved in the difference
        same = (got_idx == idx).all()
        if same:
            po, ro, f1o, no = MN.accurate_metrics(test_dict, idx, k)
            assert np.allclose([p, r, f1, n], [po, ro, f1o, no], atol=1e-5)


def test_find_lambda_fusion_sweep_matches_reference_pipeline(dev, tmp_path):
    """N4: findLambda.py with the fusion recommender — per lambda F_new = G_score * (A . HybridS), filtered top-k,
    six metrics — against the oracle pipeline; writes the reference's CSV."""
    import _stub_const
    import pandas as pd

    _stub_const.install()
    import _metrics_numpy as MN
    from _parity import assert_topk_parity
    from lgcnhs_b200.find_lambda import find_lambda
    from lgcnhs_b200.synth import synth_shape
    from model.LightGCN.model import LightGCN
    from oracle import lightgcn_oracle as LO

    d = synth_shape("small")
    tr, va, te = d.split()
    U, M, k = d.n_users, d.n_items, 10
    df = lambda s: pd.DataFrame({"user_id": d.users[s], "item_id": d.items[s]})  # noqa: E731
    torch.manual_seed(42)
    model = LightGCN(U, M, 64, 3)
    lams = [0.0, 0.5, 1.0]
    lists = []
    frame = find_lambda(U, M, df(tr), df(va), df(te), k, model=model, lambdas=lams, save_dir=str(tmp_path) + "/",
                        lists_out=lists)
    assert list(frame.columns) == ["lambda", "precision", "recall", "f1", "ndcg", "H", "I"]
    assert (tmp_path / f"lambda_evaluation_{k}.csv").exists()
    tv = np.r_[tr, va]
    A = S.interaction_matrix(U, M, d.users[tv], d.items[tv])
    Gm = S.get_spreading_general_mat(A)
    uw, iw = model.users_emb.weight.detach().cpu(), model.items_emb.weight.detach().cpu()
    e_tv = torch.from_numpy(np.stack([d.users[tv], d.items[tv]]))
    Gs = LO.masked_score(uw, iw, e_tv).numpy()
    test_dict = {}
    for u, i in zip(d.users[te].tolist(), d.items[te].tolist()):
        test_dict.setdefault(u, []).append(i)
    deg = {i: int(c) for i, c in enumerate(A.sum(0)) if c > 0}
    for n, lam in enumerate(lams):
        F_new = S.fused_resource(Gs, S.get_resource(A, S.hybrids(A, Gm, lam)))
        idx, _ = S.recommend_fast(F_new, A, k)
        got_idx = lists[n].cpu().numpy()
        # lists: tie-aware against the oracle (F_new has many exact zeros and fp32-vs-fp64 near-ties)
        assert_topk_parity(got_idx, idx, F_new, f"fusion sweep lists lambda={lam}", seen_mask=A > 0, min_checked=0.1,
                           tol_mult=4.0)
        # metrics: the device's numbers for its own lists vs the independent NumPy evaluation, +-1e-5
        p, r, f1, nd = MN.accurate_metrics(test_dict, got_idx, k)
        H, I = MN.diversity_metrics(got_idx, deg, A, k)
        got = frame.iloc[n][["precision", "recall", "f1", "ndcg", "H", "I"]].to_numpy(dtype=float)
        assert np.allclose(got, [p, r, f1, nd, H, I], atol=1e-5), (lam, got, [p, r, f1, nd, H, I])


def test_metrics_dropin_modules_match_reference_golden(dev):
    """The drop-in metrics/accurate.py + metrics/diversity.py (device only, same call signatures as the reference:
    dicts, a CPU tensor of lists, the dense float64 interaction matrix) vs the reference's recorded outputs."""
    import _stub_const
    import pandas as pd

    _stub_const.install()
    from metrics import accurate, diversity
    from utils import trans

    z = np.load(os.path.join(G, "metrics_small.npz"))
    users, items = z["users"], z["items"]
    te, tv = z["test"], np.r_[z["train"], z["val"]]
    test_dict = trans.getUserItemsDictByDataframe(pd.DataFrame({"user_id": users[te], "item_id": items[te]}))
    tv_dict = trans.getUserItemsDictByDataframe(pd.DataFrame({"user_id": users[tv], "item_id": items[tv]}))
    deg = trans.getItemDegreeByUserPosItemDict(tv_dict)
    A = trans.getInteractionMatrixByDataframe(300, 500, pd.DataFrame({"user_id": users[tv], "item_id": items[tv]}))
    rec = torch.from_numpy(z["rec"])
    acc = accurate.getAccurateMetrics(test_dict, rec, 10)
    assert np.allclose(acc, z["accurate"], atol=1e-5)
    assert np.allclose(accurate.calPrecisionAndRecall(test_dict, rec, 10), z["accurate"][:2], atol=1e-5)
    assert abs(accurate.calNDCG(test_dict, rec, 10) - z["accurate"][3]) <= 1e-5
    assert abs(accurate.calF1Score(acc[0], acc[1]) - z["accurate"][2]) <= 1e-5
    assert np.allclose(diversity.getDiversityMetrics(rec, deg, A, 10), z["diversity"], atol=1e-5)
    assert abs(diversity.calHammingDistance(rec, 10) - z["diversity"][0]) <= 1e-5
    assert abs(diversity.calInternalSimilarity(rec, deg, A, 10) - z["diversity"][1]) <= 1e-5


def test_multi_k_evaluation_matches_per_call_metrics(dev, tmp_path):
    """evaluationMetrics.py:43-96 as one batched call (lgcnhs_b200.evaluate_lists): several (model, k) list sets — dict and
    array inputs, duplicated test rows — against the independent NumPy formulas (the reference's, pinned to its outputs in
    test_cpu_host_logic.py) and against the drop-in metrics modules called the way evaluationMetrics.py calls them."""
    import _stub_const

    _stub_const.install()
    import pandas as pd

    import _metrics_numpy as MN
    from lgcnhs_b200.evaluate_lists import evaluate_lists
    from lgcnhs_b200.synth import synth_shape
    from metrics.accurate import getAccurateMetrics
    from metrics.diversity import getDiversityMetrics

    d = synth_shape("small")
    tr, va, te = d.split()
    U, M = d.n_users, d.n_items
    frame = lambda idx: pd.DataFrame({"user_id": d.users[idx], "item_id": d.items[idx]})  # noqa: E731
    train_df, val_df = frame(tr), frame(va)
    test_df = pd.concat([frame(te), frame(te[:50])])          # duplicated test rows: the reference keeps them (len(items))
    gen = np.random.default_rng(5)
    lists = {}
    for name in ("HybridS", "LightGCN"):
        for k in (5, 10, 30):
            rec = np.stack([gen.choice(M, size=k, replace=False) for _ in range(U)])
            lists[(name, k)] = {u: rec[u].tolist() for u in range(U)} if name == "HybridS" else rec
    frames = evaluate_lists(U, M, train_df, val_df, test_df, lists, save_path=str(tmp_path) + "/")
    assert sorted(frames) == [5, 10, 30] and all(list(f["Model"]) == ["HybridS", "LightGCN"] for f in frames.values())
    tv = np.r_[tr, va]
    A = S.interaction_matrix(U, M, d.users[tv], d.items[tv])
    deg = {i: int(c) for i, c in enumerate(A.sum(0)) if c > 0}
    test_dict = {}
    for u, i in zip(test_df["user_id"].tolist(), test_df["item_id"].tolist()):
        test_dict.setdefault(u, []).append(i)
    for (name, k), val in lists.items():
        rec = np.asarray([val[u] for u in range(U)]) if isinstance(val, dict) else val
        row = frames[k][frames[k]["Model"] == name].iloc[0]
        got = [row["P"], row["R"], row["F1"], row["NDCG"], row["H"], row["I"]]
        p, r, f1, n = MN.accurate_metrics(test_dict, rec, k)
        H, I = MN.diversity_metrics(rec, deg, A, k)
        assert np.allclose(got, [p, r, f1, n, H, I], atol=1e-5), (name, k, got, [p, r, f1, n, H, I])
        rt = torch.from_numpy(rec)
        assert np.allclose(got[:4], getAccurateMetrics(test_dict, rt, k), atol=1e-12)
        assert np.allclose(got[4:], getDiversityMetrics(rt, deg, A, k), atol=1e-12)
    assert os.path.exists(str(tmp_path) + "/model_evaluation_results_30.csv")
