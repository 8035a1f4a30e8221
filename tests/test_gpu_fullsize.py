"""GPU parity at BASELINE.json's FULL sizes through size-independent properties (the CPU oracle cannot
run these shapes in seconds): known answers, symmetry, linearity, conservation laws and selection
invariants that follow from the reference's formulas.

Shapes: propagation on the ML-20M-shape train graph (config 5, nnz = 32 M) and Amazon-Book shape
(config 4); spreading + top-20 on the ML-1M shape (config 2); fused LightGCN score/top-k on the
Amazon-Book shape (config 4, 52 643 x 91 599 never materialised as a whole)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _graph(dev, shape):
    import bench
    from lgcnhs_b200 import ops

    d = bench.load_shape(shape)
    adj_np, (tr, va, te) = bench.train_adj(d)
    n = d.n_users + d.n_items
    return d, ops.NormGraph(torch.from_numpy(adj_np).to(dev), n), n, (tr, va, te)


@pytest.mark.parametrize("shape", ["amazon-book", "ml-20m"])
def test_propagation_properties_fullsize(dev, shape):
    d, g, n, _ = _graph(dev, shape)
    deg = (g.rowptr[1:] - g.rowptr[:-1]).float()
    assert int(g.rowptr[-1]) == g.nnz and int(deg.sum()) == g.nnz
    # (1) known answer: A_hat D^1/2 1 = D^-1/2 A 1 = D^1/2 1, for every node, at any size
    x = deg.sqrt()[:, None].repeat(1, 64).contiguous()
    y = g.spmm(x)
    assert torch.allclose(y, x, rtol=2e-5, atol=0), f"max rel err {((y - x).abs() / x.clamp_min(1)).max():.3e}"
    # ... and it is a fixed point of the fused K-layer mean as well
    assert torch.allclose(g.propagate_mean(x, 3), x, rtol=3e-5, atol=0)
    # (2) A_hat is symmetric: <A x, y> == <x, A y>
    gen = torch.Generator(device=dev).manual_seed(0)
    a = torch.randn(n, 64, device=dev, generator=gen)
    b = torch.randn(n, 64, device=dev, generator=gen)
    lhs = (g.spmm(a).double() * b.double()).sum()
    rhs = (a.double() * g.spmm(b).double()).sum()
    assert abs(lhs - rhs) <= 1e-6 * (abs(lhs) + abs(rhs)) + 1e-3
    # (3) linearity, and the fused epilogue: alpha (A x + beta x0)
    lin = g.spmm(2.0 * a - 3.0 * b)
    ref = 2.0 * g.spmm(a) - 3.0 * g.spmm(b)
    assert (lin - ref).abs().max() <= 1e-5 * ref.abs().max()
    fused = g.spmm(a, b, alpha=0.25, beta=1.0)
    assert (fused - 0.25 * (g.spmm(a) + b)).abs().max() <= 1e-6 * fused.abs().max()
    # (4) Horner mean == explicit mean of the K+1 layer tensors
    l1 = g.spmm(a); l2 = g.spmm(l1); l3 = g.spmm(l2)
    mean = (a + l1 + l2 + l3) / 4
    assert (g.propagate_mean(a, 3) - mean).abs().max() <= 1e-5 * mean.abs().max()
    # (5) determinism: bit-identical relaunch, and row-range launches tile the full result
    assert torch.equal(l1, g.spmm(a))
    out = torch.zeros_like(a)
    cuts = [0, n // 5, n // 2, n]
    for r0, r1 in zip(cuts[:-1], cuts[1:]):
        g.spmm(a, out=out, row_begin=r0, row_end=r1)
    # a row's summation order depends on the path the launch picks for it (warp-per-row threshold sized to the
    # launch), so range launches equal the full launch to fp32 round-off, and bit for bit at a pinned threshold
    assert (out - l1).abs().max() <= 2e-6 * l1.abs().max()
    from lgcnhs_b200._lib import lib
    lib().lgc_spmm_long_row(256)
    try:
        full = g.spmm(a)
        for r0, r1 in zip(cuts[:-1], cuts[1:]):
            g.spmm(a, out=out, row_begin=r0, row_end=r1)
        assert torch.equal(out, full)
    finally:
        lib().lgc_spmm_long_row(0)
    # (6) spectral bound: |A_hat x|_2 <= |x|_2 (eigenvalues of D^-1/2 A D^-1/2 lie in [-1, 1])
    assert l1.double().norm() <= a.double().norm() * (1 + 1e-6)


def test_spreading_properties_ml1m(dev):
    import bench
    from lgcnhs_b200 import ops

    d = bench.load_shape("ml-1m")
    tr, va, _ = d.split()
    sel = np.concatenate([tr, va])
    users, items = torch.from_numpy(d.users[sel]).to(dev), torch.from_numpy(d.items[sel]).to(dev)
    eng = ops.SpreadingEngine(d.n_users, d.n_items, users, items)
    U, M = d.n_users, d.n_items
    assert int(eng.ku.sum()) == sel.size == int(eng.ki.sum())                # integer work: exact degrees
    G = eng.general_w()
    assert torch.equal(G, G.T)                                               # exact int8 path: bitwise symmetric
    # column sums: sum_i G[i,j] = sum_u A[u,j] (sum_i A[u,i]) / k_u = k_j
    assert torch.allclose(G.double().sum(0), eng.ki.double(), rtol=1e-6)
    # trace: sum_i G[i,i] = sum_u k_u / k_u = number of active users
    assert abs(G.double().diagonal().sum().item() - int((eng.ku > 0).sum())) < 1e-6 * U
    for lam in (0.0, 0.37, 1.0):
        W = eng.scale(lam, want_w32=True)
        F = eng.resource()
        if lam == 0.0:     # HeatS limit: rows of W sum to 1 -> every user's resource is conserved
            assert torch.allclose(F.double().sum(1), eng.ku.double(), rtol=1e-5)
        if lam == 1.0:     # ProbS limit: W[i,j] = G[i,j]/k_j -> columns of W sum to 1
            assert torch.allclose(W.double().sum(0)[eng.ki > 0], torch.ones((), dtype=torch.float64, device=dev), rtol=1e-5)
        # F == A @ W evaluated sparsely in float64 for a sample of users
        us = torch.arange(0, U, 97, device=dev)
        ref = torch.stack([W.double()[eng.excl_items(int(u))].sum(0) for u in us])
        assert (F[us].double() - ref).abs().max() <= 1e-5 * ref.abs().max()
        assert ((F[us].double() - ref).abs() <= 1e-5 * ref.abs() + 1e-7 * ref.abs().max()).all()
        # top-20 invariants: sorted, values are F at the ids, no seen item, nothing unseen beats the 20th
        idx, val = ops.topk_rows(F, 20, eng.excl)
        assert (val[:, :-1] >= val[:, 1:]).all() and torch.equal(torch.gather(F, 1, idx), val)
        Fm = F.clone()
        Fm[users.long(), items.long()] = -float("inf")
        assert torch.isfinite(torch.gather(Fm, 1, idx)).all()
        assert ((Fm > val[:, -1:]).sum(1) < 20).all() and ((Fm >= val[:, -1:]).sum(1) >= 20).all()


def test_lightgcn_score_topk_amazon_book(dev):
    """Full-rank eval of config 4: 52 643 users x 91 599 items, never materialised as a whole."""
    import bench
    import _stub_const
    from lgcnhs_b200 import ops

    _stub_const.install()
    from model.LightGCN.evaluation import _topk_layer0
    from model.LightGCN.model import LightGCN

    d = bench.load_shape("amazon-book")
    tr, va, te = d.split()
    torch.manual_seed(42)
    m = LightGCN(d.n_users, d.n_items, 64, 3).to(dev)
    e_tr = torch.from_numpy(np.stack([d.users[tr], d.items[tr]]))
    rec = _topk_layer0(m, d.n_users, d.n_items, [e_tr], 20)
    assert rec.shape == (d.n_users, 20) and int(rec.min()) >= 0 and int(rec.max()) < d.n_items
    xu, xi = m.users_emb.weight.detach(), m.items_emb.weight.detach()
    us = torch.arange(0, d.n_users, 1013, device=dev)
    score = (xu[us].double() @ xi.double().T).float()
    seen = ops.seen_csr(e_tr[0].to(dev), e_tr[1].to(dev), d.n_users, d.n_items)
    for r, u in enumerate(us.tolist()):
        score[r, seen[1][seen[0][u]:seen[0][u + 1]].long()] = -1024.0
    rv, ri = torch.topk(score, 20)
    got_v = torch.gather(score, 1, rec[us])
    assert torch.allclose(got_v, rv, rtol=1e-5, atol=1e-6)            # score at rank
    assert (rec[us] == ri).float().mean() > 0.999                     # ids identical except at float ties
