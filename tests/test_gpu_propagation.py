"""GPU parity: (P1-P3) CSR build, SpMM layer, fused K-layer mean vs. the CPU oracle.

Tolerance (north_star): <= 1e-5 relative, stated as
    |x - ref| <= 1e-5*|ref| + 1e-7*max|ref|   elementwise  (SURVEY.md §8d).
"""
import numpy as np
import pytest
import torch

from oracle import lightgcn_oracle as O

pytestmark = pytest.mark.gpu

RTOL, ATOL_REL = 1e-5, 1e-7


def assert_close(x, ref, what="", sum_abs=None, n_ops=8):
    """|x - ref| <= 1e-5 |ref| + 1e-7 max|ref|; for long fp32 sums (hub rows with ~10^3 terms)
    two valid summation orders may differ by ~u * sum|terms|, so when the caller passes the
    sum of absolute terms the bound also admits n_ops * 2^-24 * sum_abs (still far tighter
    than 1e-5 of the summands)."""
    x = x.detach().cpu().double()
    ref = ref.detach().cpu().double()
    tol = RTOL * ref.abs() + ATOL_REL * ref.abs().max()
    if sum_abs is not None:
        tol = tol + n_ops * 2.0 ** -24 * sum_abs.detach().cpu().double()
    bad = (x - ref).abs() > tol
    assert not bad.any(), f"{what}: {int(bad.sum())} / {bad.numel()} out of tolerance, max err {(x-ref).abs().max():.3e}"


def make_graph(name, hub=0):
    from lgcnhs_b200.synth import synth_shape, bipartite_adj

    d = synth_shape(name)
    users, items = d.users, d.items
    if hub:  # one item connected to `hub` users -> exercises the multi-chunk long-row path
        hu = np.arange(min(hub, d.n_users))
        users = np.concatenate([users, hu])
        items = np.concatenate([items, np.zeros_like(hu)])
    adj = torch.from_numpy(bipartite_adj(d.n_users, users, items))
    return d, adj


@pytest.mark.parametrize("name,hub", [("tiny", 0), ("small", 300), ("ml-100k", 0), ("ml-100k", 943)])
def test_csr_matches_gcn_norm(dev, name, hub):
    from lgcnhs_b200.ops import NormGraph

    d, adj = make_graph(name, hub)
    n = d.n_users + d.n_items
    g = NormGraph(adj.to(dev), n)
    ei, norm = O.gcn_norm(adj)
    # oracle CSR keyed by target (col), sources ascending
    order = np.lexsort((adj[0].numpy(), adj[1].numpy()))
    rowptr = np.zeros(n + 1, dtype=np.int64)
    np.add.at(rowptr, adj[1].numpy() + 1, 1)
    rowptr = np.cumsum(rowptr)
    assert np.array_equal(g.rowptr.cpu().numpy(), rowptr)                       # integer work: bit-exact
    assert np.array_equal(g.colidx.cpu().numpy()[: adj.shape[1]], adj[0].numpy()[order])
    assert np.array_equal(g.val.cpu().numpy()[: adj.shape[1]], norm.numpy()[order])  # fp32 bit-exact
    deg = (rowptr[1:] - rowptr[:-1]).astype(np.float32)
    with np.errstate(divide="ignore"):
        dinv = np.where(deg > 0, 1.0 / np.sqrt(deg), 0).astype(np.float32)
    assert np.array_equal(g.dinv.cpu().numpy(), dinv)


@pytest.mark.parametrize("name,hub,dim", [("tiny", 0, 64), ("small", 300, 64), ("ml-100k", 943, 64),
                                          ("small", 300, 32), ("small", 300, 128)])
def test_spmm_layer(dev, name, hub, dim):
    from lgcnhs_b200.ops import NormGraph

    d, adj = make_graph(name, hub)
    n = d.n_users + d.n_items
    g = NormGraph(adj.to(dev), n)
    torch.manual_seed(42)
    x = torch.randn(n, dim) * 0.1
    x0 = torch.randn(n, dim) * 0.1
    ei, norm = O.gcn_norm(adj)
    ref = O.propagate(ei, x, norm)
    sabs = O.propagate(ei, x.abs(), norm)          # sum of |terms| per output element
    y = g.spmm(x.to(dev))
    assert_close(y, ref, "A x", sum_abs=sabs)
    y2 = g.spmm(x.to(dev), x0.to(dev), alpha=0.25, beta=1.0)
    assert_close(y2, 0.25 * (ref + x0), "alpha (A x + beta x0)", sum_abs=0.25 * (sabs + x0.abs()))
    # determinism: bit-identical on a second launch (no float atomics)
    assert torch.equal(y, g.spmm(x.to(dev)))
    # row-range launch == the same rows of the full launch
    r0, r1 = n // 3, 2 * n // 3
    part = torch.zeros_like(y)
    g.spmm(x.to(dev), out=part, row_begin=r0, row_end=r1)
    assert torch.equal(part[r0:r1], y[r0:r1]) and not part[:r0].any() and not part[r1:].any()


@pytest.mark.parametrize("name,layers", [("tiny", 3), ("ml-100k", 3), ("small", 1), ("small", 0), ("small", 4)])
def test_propagate_mean(dev, name, layers):
    from lgcnhs_b200.ops import NormGraph

    d, adj = make_graph(name, hub=200)
    n = d.n_users + d.n_items
    torch.manual_seed(42)
    uw = torch.empty(d.n_users, 64).normal_(std=0.1)
    iw = torch.empty(d.n_items, 64).normal_(std=0.1)
    uf, _, itf, _ = O.lightgcn_forward(uw, iw, adj, layers)
    g = NormGraph(adj.to(dev), n)
    E = g.propagate_mean(torch.cat([uw, iw]).to(dev), layers)
    sabs = torch.stack(O.propagate_layers(torch.cat([uw, iw]).abs(), adj, layers), 0).mean(0)
    assert_close(E, torch.cat([uf, itf]), f"mean of {layers} layers", sum_abs=sabs, n_ops=8 * (layers + 1))


def test_empty_and_isolated_nodes(dev):
    """Ragged input: nodes without edges (zero degree -> dinv 0, output row 0) and an edgeless graph."""
    from lgcnhs_b200.ops import NormGraph

    adj = torch.tensor([[0, 5], [5, 0]], dtype=torch.int64)
    g = NormGraph(adj.to(dev), 8)
    x = torch.arange(8 * 64, dtype=torch.float32).reshape(8, 64)
    y = g.spmm(x.to(dev)).cpu()
    assert torch.equal(y[0], x[5]) and torch.equal(y[5], x[0]) and not y[[1, 2, 3, 4, 6, 7]].any()
    g0 = NormGraph(torch.zeros((2, 0), dtype=torch.int64, device=dev), 8)
    assert not g0.spmm(x.to(dev)).any()


def test_bad_edges_fail_loudly(dev):
    from lgcnhs_b200._lib import LgcnhsError
    from lgcnhs_b200.ops import NormGraph

    with pytest.raises(LgcnhsError):
        NormGraph(torch.tensor([[0, 9], [9, 0]], dtype=torch.int64, device=dev), 8)
    with pytest.raises(LgcnhsError):
        NormGraph(torch.tensor([[0, 1], [1, 0]], dtype=torch.int64), 8)  # CPU tensor: no fallback


def test_backward_on_non_symmetric_graph_matches_autograd(dev):
    """the backward graph must be the STRUCTURAL transpose of A_hat with the forward values; re-running
    gcn_norm on the transposed edge list (out-degree normalisation) is only equal for symmetric graphs.  A directed
    random graph, with a hub target so that the chunk lists of the transpose are exercised, vs autograd on the oracle."""
    from lgcnhs_b200.propagation import PropagateMean, graphs_for

    g = np.random.default_rng(2)
    n, e = 700, 9000
    src, dst = g.integers(0, n, e), g.integers(0, n, e)
    src = np.r_[src, np.full(600, 5)]                  # node 5 fans OUT to 600 targets -> long row of the transpose
    dst = np.r_[dst, g.permutation(n)[:600]]
    key = np.unique(src * n + dst)
    ei = torch.from_numpy(np.stack([key // n, key % n]))
    fwd, bwd = graphs_for(ei.to(dev), n)
    assert bwd is not fwd and bwd.n_chunks >= 1
    torch.manual_seed(0)
    x = torch.randn(n, 64) * 0.1
    w = torch.randn(n, 64)
    xr = x.clone().requires_grad_()
    layers = O.propagate_layers(xr, ei, 3)
    ref = sum(layers) / 4
    (ref * w).sum().backward()
    xd = x.to(dev).requires_grad_()
    out = PropagateMean.apply(xd, fwd, bwd, 3)
    assert_close(out, ref.detach(), "forward on a directed graph")
    (out * w.to(dev)).sum().backward()
    absA = sum(O.propagate_layers(w.abs(), torch.stack([ei[1], ei[0]]), 3)) / 4
    assert_close(xd.grad, xr.grad, "dX0 on a directed graph", sum_abs=absA + 1e-6)


def test_pipelined_host_table_propagation(dev):
    """PipelinedPropagation: upload / K layers / download on three streams with two buffer sets — every call must return
    exactly what a plain propagate_mean of ITS input gives, also when calls with different inputs are in flight."""
    from lgcnhs_b200.ops import NormGraph
    from lgcnhs_b200.propagation import PipelinedPropagation

    d, adj = make_graph("ml-100k")
    n = d.n_users + d.n_items
    adj = adj.to(dev)
    g = NormGraph(adj, n)
    pipe = PipelinedPropagation(adj, n, 64, 3)
    gen = torch.Generator().manual_seed(1)
    ins = [torch.randn(n, 64, generator=gen).pin_memory() for _ in range(5)]
    outs = [torch.empty(n, 64).pin_memory() for _ in range(5)]
    for x, o in zip(ins, outs):
        pipe.submit(x, o)
    pipe.synchronize()
    for x, o in zip(ins, outs):
        assert torch.equal(o, g.propagate_mean(x.to(dev), 3).cpu())
    with pytest.raises(RuntimeError):
        pipe.submit(torch.randn(n, 64), outs[0])          # unpinned host memory is refused


@pytest.mark.parametrize("name,hub,dim,layers", [("tiny", 0, 64, 3), ("small", 300, 64, 3), ("ml-100k", 0, 64, 3),
                                                 ("ml-100k", 943, 32, 2), ("ml-100k", 943, 128, 1), ("douban", 0, 64, 4)])
def test_cooperative_k_layer_kernel(dev, name, hub, dim, layers):
    """lgc_propagate_mean_coop (all layers in one cooperative launch, rows cut into <= 128-nnz units, long rows combined
    from partials in unit order) vs the oracle's per-layer tensors and vs the per-layer launches."""
    from lgcnhs_b200.ops import NormGraph

    d, adj = make_graph(name, hub)
    n = d.n_users + d.n_items
    g = NormGraph(adj.to(dev), n)
    torch.manual_seed(3)
    x0 = torch.randn(n, dim) * 0.1
    ref = sum(O.propagate_layers(x0, adj, layers)) / (layers + 1)
    absref = sum(O.propagate_layers(x0.abs(), adj, layers)) / (layers + 1)
    got = g.propagate_mean(x0.to(dev), layers, coop=True)
    assert_close(got, ref, f"cooperative K-layer mean ({name})", sum_abs=absref)
    plain = g.propagate_mean(x0.to(dev), layers, coop=False)
    assert_close(got, plain, "cooperative vs per-layer launches", sum_abs=absref)
    assert torch.equal(got, g.propagate_mean(x0.to(dev), layers, coop=True))          # deterministic
    cu = g.coop_units()
    assert int((cu["end"] - cu["start"]).max()) <= 128 and int((cu["end"] - cu["start"]).sum()) == g.nnz
    if hub:
        assert cu["n_split"] >= 1


@pytest.mark.parametrize("name,hub,dim,frac", [("tiny", 0, 64, 0.2), ("small", 300, 64, 0.02), ("ml-100k", 943, 64, 0.02),
                                               ("ml-100k", 943, 32, 0.05), ("ml-100k", 0, 64, 0.0), ("ml-100k", 0, 64, 1.0)])
def test_masked_source_layer_is_bit_identical(dev, name, hub, dim, frac):
    """Sparse-source SpMM (lgc_spmm_layer_masked / lgc_propagate_mean_masked): with X zero outside the masked rows the
    result must be the one of the unmasked call, bit for bit — the skipped terms are exact zeros.  Also checked against
    the oracle layer, and on the long-row (chunk) path through the hub item."""
    from lgcnhs_b200 import ops
    from lgcnhs_b200.ops import NormGraph

    d, adj = make_graph(name, hub)
    n = d.n_users + d.n_items
    g = NormGraph(adj.to(dev), n)
    gen = torch.Generator().manual_seed(7)
    live = torch.rand(n, generator=gen) < frac
    if hub:
        live[d.n_users] = True                                  # the hub item itself is a live source of the user rows
        live[: min(hub, d.n_users) : 7] = True                  # and some of the hub's sources are live
    X = torch.randn(n, dim, generator=gen) * live[:, None].float()
    mask = ops.row_mask_words(n, dev)
    rows = torch.nonzero(live).flatten()
    words = torch.zeros(mask.numel(), dtype=torch.int64)
    words.index_add_(0, rows // 32, torch.ones_like(rows) << (rows % 32))
    mask.copy_(torch.where(words >= 2 ** 31, words - 2 ** 32, words).to(torch.int32))
    Xd = X.to(dev)
    plain = g.spmm(Xd, Xd, 0.25, 1.0)
    masked = g.spmm(Xd, Xd, 0.25, 1.0, src_mask=mask)
    assert torch.equal(plain, masked)
    out_p = g.propagate_mean(Xd, 3)
    out_m = g.propagate_mean(Xd, 3, x0_row_mask=mask)
    assert torch.equal(out_p, out_m)
    # a row range (the multi-GPU partition's launch shape) through the row-list entry
    a, b = n // 3, n - n // 5
    y0 = torch.zeros_like(Xd)
    y1 = torch.zeros_like(Xd)
    g.spmm_rows_bcast(Xd, Xd, 1.0, 1.0, [y0.data_ptr()], [(a, b)])
    g.spmm_rows_bcast(Xd, Xd, 1.0, 1.0, [y1.data_ptr()], [(a, b)], src_mask=mask)
    assert torch.equal(y0, y1)


def test_row_mask_batch_sets_and_clears(dev):
    from lgcnhs_b200 import ops

    U, M, B = 1000, 777, 300
    gen = torch.Generator().manual_seed(3)
    users = torch.randint(0, U, (B,), generator=gen).to(dev)
    pos = torch.randint(0, M, (B,), generator=gen).to(dev)
    neg = torch.randint(0, M, (B,), generator=gen).to(dev)
    mask = ops.row_mask_words(U + M, dev)
    ops.row_mask_batch(mask, users, pos, neg, U, True)
    bits = ((mask.cpu().numpy().view(np.uint32)[:, None] >> np.arange(32, dtype=np.uint32)[None, :]) & 1).reshape(-1)[: U + M]
    want = np.zeros(U + M, dtype=np.uint32)
    want[users.cpu().numpy()] = 1
    want[U + pos.cpu().numpy()] = 1
    want[U + neg.cpu().numpy()] = 1
    assert np.array_equal(bits, want)
    ops.row_mask_batch(mask, users, pos, neg, U, False)
    assert int(mask.abs().sum()) == 0


def test_masked_source_rejects_dim_128(dev):
    from lgcnhs_b200 import ops
    from lgcnhs_b200._lib import LgcnhsError
    from lgcnhs_b200.ops import NormGraph

    d, adj = make_graph("tiny")
    n = d.n_users + d.n_items
    g = NormGraph(adj.to(dev), n)
    X = torch.zeros(n, 128, device=dev)
    with pytest.raises(LgcnhsError):
        g.spmm(X, X, 1.0, 1.0, src_mask=ops.row_mask_words(n, dev))
