"""CPU: the oracle restatement against the golden vectors recorded from the reference's own code
(oracle/make_golden.py).  Integer/index results bit-exact; fp64 NumPy results bit-exact on the
same BLAS, else within 1e-12 relative (summation order inside np.dot)."""
import os

import numpy as np
import pytest
import torch

from oracle import lightgcn_oracle as LO
from oracle import spread_oracle as SO

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def close64(a, b):
    return np.allclose(a, b, rtol=1e-12, atol=0)


@pytest.mark.parametrize("name", ["tiny", "small"])
def test_spreading_oracle_matches_reference(name):
    z = np.load(os.path.join(G, f"spread_{name}.npz"))
    users, items = z["users"], z["items"]
    U, M = {"tiny": (96, 160), "small": (300, 500)}[name]
    sel = np.r_[z["train"], z["val"]]
    A = SO.interaction_matrix(U, M, users[sel], items[sel])
    Gm = SO.get_spreading_general_mat(A)
    if "G" in z:
        assert close64(Gm, z["G"])
    k = int(z["k"])
    seen = {}
    for u, i in zip(users[sel].tolist(), items[sel].tolist()):
        seen.setdefault(u, []).append(i)
    for key in [f for f in z.files if f.startswith("F_")]:
        lam = float(key[2:])
        W = SO.hybrids(A, Gm, lam)
        F = SO.get_resource(A, W)
        if f"W_{lam}" in z:
            assert close64(W, z[f"W_{lam}"])
        assert close64(F, z[key])
        # the reference loop on the reference's F: integer result, bit-exact
        rec = SO.recommend_loop(z[key], seen, k)
        assert np.array_equal(np.array([rec[u] for u in range(U)]), z[f"rec_{lam}"])
        # vectorised rule == loop wherever scores at rank are distinct
        fi, fv = SO.recommend_fast(z[key], A, k)
        distinct = np.r_["1", np.ones((U, 1), bool), fv[:, 1:] != fv[:, :-1]] & np.r_["1", fv[:, 1:] != fv[:, :-1], np.ones((U, 1), bool)]
        assert np.array_equal(fi[distinct], z[f"rec_{lam}"][distinct])


def test_spreading_properties():
    """Known-answer properties: mass conservation at lambda=0, finite W for zero-degree items."""
    rng = np.random.default_rng(1)
    A = (rng.random((50, 40)) < 0.2).astype(np.float64)
    A[:, 7] = 0           # zero-degree item
    A[3, :] = 0           # zero-degree user
    Gm = SO.get_spreading_general_mat(A)
    for lam in (0.0, 0.5, 1.0):
        W = SO.hybrids(A, Gm, lam)
        assert np.isfinite(W).all() and not W[7].any() and not W[:, 7].any()
    F = SO.get_resource(A, SO.hybrids(A, Gm, 0.0))
    assert np.allclose(F.sum(1), A.sum(1))
    assert np.allclose(SO.hybrids(A, Gm, 1.0), SO.probs(A, Gm)) and np.allclose(SO.hybrids(A, Gm, 0.0), SO.heats(A, Gm))


def test_lightgcn_oracle_matches_reference():
    z = np.load(os.path.join(G, "lightgcn_tiny.npz"))
    U, M = 96, 160
    ei = torch.from_numpy(np.stack([z["users"][z["train"]], z["items"][z["train"]]]))
    adj = LO.convert_edge_index_to_adj(U, M, ei)
    assert np.array_equal(adj.numpy(), z["adj"])                                       # index work: bit-exact
    assert np.array_equal(LO.convert_adj_to_edge_index(U, M, adj).numpy(), z["edge_back"])
    uw = torch.from_numpy(z["users_w"]).requires_grad_()
    iw = torch.from_numpy(z["items_w"]).requires_grad_()
    uf, u0, itf, i0 = LO.lightgcn_forward(uw, iw, adj, 3)
    assert np.array_equal(uf.detach().numpy(), z["users_final"]) and np.array_equal(itf.detach().numpy(), z["items_final"])
    u, p, n = (torch.from_numpy(z[k]) for k in ("bpr_u", "bpr_p", "bpr_n"))
    loss = LO.bpr_loss(uf[u], u0[u], itf[p], i0[p], itf[n], i0[n], 1e-6)
    assert loss.item() == float(z["bpr_loss"])
    loss.backward()
    assert np.allclose(uw.grad.numpy(), z["grad_users"], rtol=1e-6, atol=1e-10)
    assert np.allclose(iw.grad.numpy(), z["grad_items"], rtol=1e-6, atol=1e-10)
    # Horner form of the layer mean (what the CUDA path computes) == stack/mean within fp32 rounding
    ei2, norm = LO.gcn_norm(adj)
    x0 = torch.cat([uw, iw]).detach()
    s = x0
    for _ in range(3):
        s = LO.propagate(ei2, s, norm) + x0
    assert torch.allclose(s / 4, torch.cat([uf, itf]).detach(), rtol=1e-5, atol=1e-8)


def test_negative_sampling_contract():
    z = np.load(os.path.join(G, "lightgcn_tiny.npz"))
    ei = torch.from_numpy(z["edge_back"])
    pos = set(zip(ei[0].tolist(), ei[1].tolist()))
    # golden sample drawn by the reference's sampleMiniBatch: never a positive pair, edges come from the graph
    for u, p, n in zip(z["sample_u"].tolist(), z["sample_p"].tolist(), z["sample_n"].tolist()):
        assert (u, p) in pos and (u, n) not in pos
    g = torch.Generator().manual_seed(3)
    u, p, n = LO.structured_negative_sampling(ei, generator=g)
    assert all((a, b) not in pos for a, b in zip(u.tolist(), n.tolist()))
    assert int(n.max()) < int(ei.max()) + 1          # range quirk: [0, max id + 1)
    u, p, n = LO.structured_negative_sampling(ei, contains_neg_self_loops=False, generator=g)
    assert (u != n).all()
