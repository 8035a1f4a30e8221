"""Generate tests/golden/*.npz by running the REFERENCE'S OWN CODE in the build container.

    python oracle/make_golden.py

* spreading path: /root/reference/model/SpreadMethod/{model,recommend}.py are imported and run
  as they are (NumPy fp64);
* LightGCN path: /root/reference/model/LightGCN/{model,loss}.py and utils/graph.py are imported on
  top of oracle/pyg_stub (PyG / torch_sparse are not installable offline);
* formats + metrics: /root/reference/utils/trans.py, metrics/{accurate,diversity}.py as they are.
Each fixture is also checked against the oracle restatement before it is written, so a committed
fixture certifies "oracle == reference on this input".  /root/reference is read-only and is never
copied; only inputs/outputs are stored.
"""
import os
import random
import sys
import tempfile

import numpy as np
import pandas as pd
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "light-graph-convolutional-recommendation-algorithm-based-on-hybrid-spreading_b200")
REF = "/root/reference"
OUT = os.path.join(ROOT, "tests", "golden")


def main():
    os.makedirs(OUT, exist_ok=True)
    scratch = tempfile.mkdtemp(prefix="lgc_golden_")
    os.chdir(scratch)                       # const.py mkdirs ./RS/... relative to the cwd
    # the drop-in tree has regular packages named model/utils/metrics, which would shadow the
    # reference's namespace packages: import what is needed from it first, then drop it from the path
    sys.path.insert(0, PKG)
    from lgcnhs_b200.synth import bipartite_adj, synth_shape
    sys.path.remove(PKG)
    sys.path[:0] = [REF, os.path.join(ROOT, "oracle", "pyg_stub"), ROOT]
    import const  # noqa: F401  (reference's, unchanged)
    from metrics import accurate as ref_acc, diversity as ref_div
    from model.LightGCN import loss as ref_loss, model as ref_model
    from model.SpreadMethod import model as ref_spread, recommend as ref_rec
    from utils import graph as ref_graph, trans as ref_trans
    assert ref_spread.__file__.startswith(REF) and ref_model.__file__.startswith(REF)

    from oracle import lightgcn_oracle as LO, spread_oracle as SO

    # ------------------------------------------------------------------ spreading (S0-S5)
    for name, k in (("tiny", 10), ("small", 20)):
        d = synth_shape(name)
        tr, va, te = d.split()
        df = lambda idx: pd.DataFrame({"user_id": d.users[idx], "item_id": d.items[idx], "rating": 5,  # noqa: E731
                                       "rating_time": "2024-01-01 00:00:00"})
        train_df, val_df, test_df = df(tr), df(va), df(te)
        both = pd.concat([train_df, val_df])
        A = ref_trans.getInteractionMatrixByDataframe(d.n_users, d.n_items, both)
        G = ref_spread.getSpreadingGeneralMat(A)
        out = {"users": d.users, "items": d.items, "train": tr, "val": va, "test": te, "k": np.array(k)}
        if name == "tiny":
            out["G"] = G
        assert np.array_equal(A, SO.interaction_matrix(d.n_users, d.n_items, both.user_id, both.item_id))
        assert np.array_equal(G, SO.get_spreading_general_mat(A))
        const.cfg.RECOMMEND["k"] = k
        const.cfg.MODEL["name"] = "HybridS"
        for lam in ((0.0, 0.3, 1.0) if name == "tiny" else (0.3,)):
            W = ref_spread.HybridS(A, G, lam)
            F = ref_spread.getResource(A, W)
            assert np.array_equal(W, SO.hybrids(A, G, lam)) and np.array_equal(F, SO.get_resource(A, W))
            rec = ref_rec.recommendForAllUser(F, d.n_users, train_df, val_df, k)
            rec = np.array([rec[u] for u in range(d.n_users)], dtype=np.int64)
            seen = ref_trans.getUserItemsDictByDataframe(both)
            loop = SO.recommend_loop(F, seen, k)
            assert all(np.array_equal(rec[u], np.asarray(loop[u])) for u in range(d.n_users))
            out[f"F_{lam}"], out[f"rec_{lam}"] = F, rec
            if name == "tiny":
                out[f"W_{lam}"] = W
        assert np.array_equal(ref_spread.ProbS(A, G), SO.probs(A, G)) and np.array_equal(ref_spread.HeatS(A, G), SO.heats(A, G))
        np.savez_compressed(os.path.join(OUT, f"spread_{name}.npz"), **out)

    # ------------------------------------------------------------------ LightGCN (P0-P6, P8)
    d = synth_shape("tiny")
    tr, va, te = d.split()
    ei = torch.from_numpy(np.stack([d.users[tr], d.items[tr]]))
    adj = ref_graph.convertEdgeIndexToAdjMatrix(d.n_users, d.n_items, ei)        # reference, dense route
    back = ref_graph.convertAdjMatrixToEdgeIndex(d.n_users, d.n_items, adj)
    assert torch.equal(adj, LO.convert_edge_index_to_adj(d.n_users, d.n_items, ei))
    assert torch.equal(back, LO.convert_adj_to_edge_index(d.n_users, d.n_items, adj))
    assert torch.equal(adj, torch.from_numpy(bipartite_adj(d.n_users, d.users[tr], d.items[tr])))
    torch.manual_seed(42)
    model = ref_model.LightGCN(d.n_users, d.n_items, 64, 3)                      # reference class on the stub
    uf, u0, itf, i0 = model.forward(adj)
    ouf, _, oitf, _ = LO.lightgcn_forward(model.users_emb.weight, model.items_emb.weight, adj, 3)
    assert torch.equal(uf, ouf) and torch.equal(itf, oitf)
    g = torch.Generator().manual_seed(7)
    B = 64
    u = torch.randint(d.n_users, (B,), generator=g)
    p = torch.randint(d.n_items, (B,), generator=g)
    n = torch.randint(d.n_items, (B,), generator=g)
    loss = ref_loss.BPRLoss(uf[u], u0[u], itf[p], i0[p], itf[n], i0[n], 1e-6)
    assert torch.equal(loss, LO.bpr_loss(uf[u], u0[u], itf[p], i0[p], itf[n], i0[n], 1e-6))
    loss.backward()
    random.seed(0)
    torch.manual_seed(0)
    su, sp, sn = ref_loss.sampleMiniBatch(32, back)
    np.savez_compressed(os.path.join(OUT, "lightgcn_tiny.npz"),
                        users=d.users, items=d.items, train=tr, val=va, test=te, adj=adj.numpy(), edge_back=back.numpy(),
                        users_w=model.users_emb.weight.detach().numpy(), items_w=model.items_emb.weight.detach().numpy(),
                        users_final=uf.detach().numpy(), items_final=itf.detach().numpy(),
                        bpr_u=u.numpy(), bpr_p=p.numpy(), bpr_n=n.numpy(), bpr_loss=loss.detach().numpy(),
                        grad_users=model.users_emb.weight.grad.numpy(), grad_items=model.items_emb.weight.grad.numpy(),
                        sample_u=su.numpy(), sample_p=sp.numpy(), sample_n=sn.numpy())

    # ------------------------------------------------------------------ config 3: LightGCNOpti + SpreadLightGCNOpti
    # The reference's own model/LightGCNOpti/model.py and model/SpreadLightGCNOpti/{model,recommend}.py, run as
    # they are on the stub.  The pickled module the reference loads (<k>_LightGCNOpti.pth) is a seeded, untrained
    # LightGCNOpti of the reference's own class; torch.load is given weights_only=False (torch >= 2.6 default
    # changed; the reference's environment.yaml pins an older torch) — nothing of the reference's logic is altered.
    from model.LightGCNOpti import model as ref_opti_model
    from model.SpreadLightGCNOpti import model as ref_slo_model, recommend as ref_slo_rec
    assert ref_opti_model.__file__.startswith(REF) and ref_slo_model.__file__.startswith(REF)
    d = synth_shape("tiny")
    tr, va, te = d.split()
    k, lam = 10, 0.3
    frng = np.random.default_rng(11)
    FU, FI = 29, 31                                   # ML-100K-like feature widths
    user_feat = np.round(frng.random((d.n_users, FU)), 3) * (frng.random((d.n_users, FU)) < 0.4)
    item_feat = np.round(frng.random((d.n_items, FI)), 3) * (frng.random((d.n_items, FI)) < 0.4)
    perm_u, perm_i = frng.permutation(d.n_users), frng.permutation(d.n_items)      # files are not sorted by id
    ufd = pd.DataFrame({"user_id": perm_u, "user_features": [str(user_feat[u].tolist()) for u in perm_u]})
    ifd = pd.DataFrame({"item_id": perm_i, "item_features": [str(item_feat[i].tolist()) for i in perm_i]})
    df = lambda idx: pd.DataFrame({"user_id": d.users[idx], "item_id": d.items[idx], "rating": 5,  # noqa: E731
                                   "rating_time": "2024-01-01 00:00:00"})
    rating_df, train_df, val_df, test_df = df(np.arange(d.users.size)), df(tr), df(va), df(te)
    torch.manual_seed(42)                             # reference LightGCNOpti/train.py:94 seeds before the constructor
    opti = ref_opti_model.LightGCNOpti(d.n_users, d.n_items, 64, 3, torch.from_numpy(user_feat).float(),
                                       torch.from_numpy(item_feat).float())
    adj = ref_graph.convertEdgeIndexToAdjMatrix(d.n_users, d.n_items, torch.from_numpy(np.stack([d.users[tr], d.items[tr]])))
    ouf, ou0, oitf, oi0 = opti.forward(adj)
    chk_u, _, chk_i, _ = LO.lightgcn_forward(opti.users_emb.weight, opti.items_emb.weight, adj, 3)
    assert torch.equal(ouf, chk_u) and torch.equal(oitf, chk_i)
    const.cfg.RECOMMEND["k"] = k
    const.cfg.MODEL["name"] = "SpreadLightGCNOpti"
    const.cfg.MODEL["HyperParameter"]["lambda"] = lam
    torch.save(opti, const.cfg.MODEL["save_path"] + str(k) + "_LightGCNOpti.pth")
    _load = torch.load
    torch.load = lambda *a, **kw: _load(*a, **{**kw, "weights_only": False})
    try:
        Gs = ref_slo_model.getAllocateMat(d.n_users, d.n_items, rating_df, train_df, val_df, test_df, ufd, ifd, k)
        F_new = ref_slo_model.getResourceMat(d.n_users, d.n_items, rating_df, train_df, val_df, test_df, ufd, ifd)
        rec = ref_slo_rec.recommendSpreadLightGCNOpti(d.n_users, d.n_items, rating_df, train_df, val_df, test_df, ufd, ifd)
    finally:
        torch.load = _load
    rec = np.array([rec[u] for u in range(d.n_users)], dtype=np.int64)
    both = pd.concat([train_df, val_df])
    A = SO.interaction_matrix(d.n_users, d.n_items, both.user_id, both.item_id)
    e_tr = torch.from_numpy(np.stack([d.users[tr], d.items[tr]]))
    e_va = torch.from_numpy(np.stack([d.users[va], d.items[va]]))
    oGs = LO.masked_score(opti.users_emb.weight.detach(), opti.items_emb.weight.detach(), e_tr, e_va).numpy()
    oF = SO.fused_resource(oGs, SO.get_resource(A, SO.hybrids(A, SO.get_spreading_general_mat(A), lam)))
    assert np.array_equal(Gs, oGs) and np.array_equal(F_new, oF)                 # oracle == reference, bit for bit
    loop = SO.recommend_loop(F_new, ref_trans.getUserItemsDictByDataframe(both), k)
    assert all(np.array_equal(rec[u], np.asarray(loop[u])) for u in range(d.n_users))
    np.savez_compressed(os.path.join(OUT, "opti_tiny.npz"), users=d.users, items=d.items, train=tr, val=va, test=te,
                        user_feat=user_feat, item_feat=item_feat, perm_u=perm_u, perm_i=perm_i,
                        users_w=opti.users_emb.weight.detach().numpy(), items_w=opti.items_emb.weight.detach().numpy(),
                        users_final=ouf.detach().numpy(), items_final=oitf.detach().numpy(), adj=adj.numpy(),
                        G_score=Gs, F_new=F_new, rec=rec, k=np.array(k), lam=np.array(lam))

    # ------------------------------------------------------------------ formats + metrics
    d = synth_shape("small")
    tr, va, te = d.split()
    k = 10
    rng = np.random.default_rng(5)
    rec = np.stack([rng.choice(d.n_items, size=k, replace=False) for _ in range(d.n_users)])
    rec[:, :3] = rec[0, :3]                                    # overlap between users
    rec_t = torch.from_numpy(rec)
    test_df = pd.DataFrame({"user_id": d.users[te], "item_id": d.items[te]})
    tv_df = pd.DataFrame({"user_id": d.users[np.r_[tr, va]], "item_id": d.items[np.r_[tr, va]]})
    test_dict = ref_trans.getUserItemsDictByDataframe(test_df)
    tv_dict = ref_trans.getUserItemsDictByDataframe(tv_df)
    deg = ref_trans.getItemDegreeByUserPosItemDict(tv_dict)
    A = ref_trans.getInteractionMatrixByDataframe(d.n_users, d.n_items, tv_df)
    acc = ref_acc.getAccurateMetrics(test_dict, rec_t, k)
    div = ref_div.getDiversityMetrics(rec_t, deg, A, k)
    np.savez_compressed(os.path.join(OUT, "metrics_small.npz"), users=d.users, items=d.items, train=tr, val=va, test=te,
                        rec=rec, accurate=np.array(acc), diversity=np.array(div),
                        dict_keys=np.array(list(test_dict.keys())), dict_first=np.array([v[0] for v in test_dict.values()]),
                        deg_items=np.array(list(deg.keys())), deg_vals=np.array(list(deg.values())))
    print("golden fixtures written to", OUT, [(f, os.path.getsize(os.path.join(OUT, f))) for f in sorted(os.listdir(OUT))])


if __name__ == "__main__":
    main()
