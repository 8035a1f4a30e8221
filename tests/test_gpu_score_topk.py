"""GPU parity: (P8 / S4) score block + seen fill, and row-wise masked top-k.

Top-k ids must be identical except at float ties (north_star), so ids are compared where the
reference scores are distinct and the score-at-rank everywhere."""
import numpy as np
import pytest
import torch

from oracle import lightgcn_oracle as O
from oracle import spread_oracle as S
from test_gpu_propagation import assert_close

pytestmark = pytest.mark.gpu


def test_score_block_and_fill(dev):
    from lgcnhs_b200 import ops
    from lgcnhs_b200.synth import synth_shape

    d = synth_shape("ml-100k")
    torch.manual_seed(42)
    uw = torch.empty(d.n_users, 64).normal_(std=0.1)
    iw = torch.empty(d.n_items, 64).normal_(std=0.1)
    tr, va, _ = d.split()
    e_tr = torch.from_numpy(np.stack([d.users[tr], d.items[tr]]))
    e_va = torch.from_numpy(np.stack([d.users[va], d.items[va]]))
    ref = O.masked_score(uw, iw, e_tr, e_va)
    seen = ops.seen_csr(torch.cat([e_tr[0], e_va[0]]).to(dev), torch.cat([e_tr[1], e_va[1]]).to(dev), d.n_users, d.n_items)
    out = ops.score_block(uw.to(dev), iw.to(dev), 0, d.n_users, seen)
    assert_close(out, ref, "masked score")
    assert (out.cpu() == -1024).sum() == (ref == -1024).sum()
    # a user sub-block with an odd offset
    blk = ops.score_block(uw.to(dev), iw.to(dev), 101, 777, seen)
    assert torch.equal(blk, out[101:777])
    # P8 end to end: torch.topk(score, k) ids
    for k in (10, 20, 100):
        idx, val = ops.topk_rows(out, k)
        rv, ri = O.topk_items(ref, k)
        assert_close(val, rv, "top-k scores")
        distinct = (rv[:, 1:] != rv[:, :-1]).all(dim=1)
        assert torch.equal(idx.cpu()[distinct], ri[distinct])


@pytest.mark.parametrize("rows,cols,k", [(7, 33, 5), (64, 1682, 20), (33, 3706, 100), (5, 26744, 20), (3, 70000, 128),
                                         (4, 20, 20)])
def test_topk_rows_random(dev, rows, cols, k):
    from lgcnhs_b200 import ops

    g = torch.Generator().manual_seed(rows * 1000 + cols)
    Sm = torch.randn(rows, cols, generator=g)
    Sm[0, : cols // 2] = 0.0          # heavy ties
    if rows > 1:
        Sm[1] = -Sm[1].abs()          # all negative
    idx, val = ops.topk_rows(Sm.to(dev), k)
    rv, _ = torch.topk(Sm, k)
    assert torch.equal(val.cpu(), rv)                       # selection + order are exact
    assert torch.equal(torch.gather(Sm, 1, idx.cpu()), rv)  # ids point at those values
    assert all(len(set(r.tolist())) == k for r in idx.cpu())


def test_topk_ties_prefer_larger_index(dev):
    from lgcnhs_b200 import ops

    Sm = torch.zeros(2, 50)
    Sm[1, 7] = 1.0
    idx, _ = ops.topk_rows(Sm.to(dev), 3)
    assert idx.cpu().tolist() == [[49, 48, 47], [7, 49, 48]]   # np.argsort(row)[::-1] order


def test_topk_with_exclusion_matches_reference_loop(dev):
    """S4: argsort + Python filter loop (model/SpreadMethod/recommend.py:35-47)."""
    from lgcnhs_b200 import ops
    from lgcnhs_b200.synth import synth_shape

    d = synth_shape("small")
    g = torch.Generator().manual_seed(3)
    F = torch.rand(d.n_users, d.n_items, generator=g)
    F[F < 0.3] = 0.0
    F = F.double()      # float64 like the reference, but exactly representable in fp32
    A = S.interaction_matrix(d.n_users, d.n_items, d.users, d.items)
    seen = {}
    for u, i in zip(d.users.tolist(), d.items.tolist()):
        seen.setdefault(u, []).append(i)
    ref = S.recommend_loop(F.numpy(), seen, 20)
    excl = ops.ExclusionMask.from_pairs(torch.from_numpy(d.users).to(dev), torch.from_numpy(d.items).to(dev), d.n_users, d.n_items)
    F32 = F.float()
    idx, val = ops.topk_rows(F32.to(dev), 20, excl)
    idx = idx.cpu().numpy()
    fi, fv = S.recommend_fast(F32.double().numpy(), A, 20)
    assert np.array_equal(idx, fi)                               # same rule, vectorised oracle
    for u in range(d.n_users):
        r = np.asarray(ref[u])
        assert not set(idx[u].tolist()) & set(seen.get(u, []))   # never recommends a seen item
        assert np.array_equal(F.numpy()[u, r].astype(np.float32), val.cpu().numpy()[u])   # score at rank
        ok = np.r_[True, fv[u][1:] != fv[u][:-1]] & np.r_[fv[u][1:] != fv[u][:-1], True]
        assert np.array_equal(idx[u][ok], r[ok])                 # ids equal wherever scores are distinct


def test_topk_refill_and_mask_offsets(dev):
    """Adversarial layout: the whole top-k sits in ONE lane's stride (columns = lane mod 32), so the
    per-lane list must refill many times; plus a mask whose rows start at unaligned bit offsets."""
    from lgcnhs_b200 import ops

    rows, cols, k = 37, 1000, 100
    Sm = torch.rand(rows, cols) * 0.01
    hot = torch.arange(5, cols, 32)[:31]            # all in lane 5's share
    Sm[:, hot] = 10.0 + torch.rand(rows, hot.numel())
    g = torch.Generator().manual_seed(1)
    ex_u = torch.randint(rows + 11, (4000,), generator=g)
    ex_i = torch.randint(cols, (4000,), generator=g)
    mask = ops.ExclusionMask.from_pairs(ex_u.to(dev), ex_i.to(dev), rows + 11, cols)
    for off in (0, 11):
        idx, val = ops.topk_rows(Sm.to(dev), k, mask, row_offset=off)
        ref = Sm.clone()
        sel = (ex_u >= off) & (ex_u < off + rows)
        ref[ex_u[sel] - off, ex_i[sel]] = -float("inf")
        rv, ri = torch.topk(ref, k)
        assert torch.equal(val.cpu(), rv) and torch.equal(torch.gather(ref, 1, idx.cpu()), rv)


# ------------------------------------------------------------------------------------------
# fused score + seen rule + top-k (lgc_score_topk): the (U, M) matrix is never materialised
# ------------------------------------------------------------------------------------------
def _rand_problem(U, M, dim, seed, n_seen):
    g = torch.Generator().manual_seed(seed)
    uw = torch.empty(U, dim).normal_(std=0.1, generator=g)
    iw = torch.empty(M, dim).normal_(std=0.1, generator=g)
    su = torch.randint(U, (n_seen,), generator=g)
    si = torch.randint(M, (n_seen,), generator=g)
    return uw, iw, su, si


@pytest.mark.parametrize("U,M,dim,k", [(943, 1682, 64, 20), (943, 1682, 64, 100), (130, 3706, 64, 10), (65, 100, 64, 32),
                                       (7, 129, 32, 5), (200, 5000, 32, 128), (64, 128, 64, 128)])
def test_score_topk_equals_unfused(dev, U, M, dim, k):
    """Same FMA order as lgc_score_block, same selection rule as lgc_topk_rows -> bit-identical ids and values;
    and the reference semantics (matmul, score[seen] = -1024, torch.topk; model/LightGCN/recommend.py:86-114)."""
    from lgcnhs_b200 import ops

    uw, iw, su, si = _rand_problem(U, M, dim, U * 7 + M, 20 * U)
    seen = ops.seen_csr(su.to(dev), si.to(dev), U, M)
    dense = ops.score_block(uw.to(dev), iw.to(dev), 0, U, seen)
    ridx, rval = ops.topk_rows(dense, k)
    idx, val = ops.score_topk(uw.to(dev), iw.to(dev), k, seen, precise=True)
    assert torch.equal(val, rval) and torch.equal(idx, ridx)
    # the CPU statement of the reference rule
    ref = uw @ iw.T
    ref[su, si] = -1024.0
    rv, _ = torch.topk(ref, k)
    assert_close(val, rv, "fused top-k scores")
    # a user sub-range with an odd offset
    if U > 70:
        idx2, val2 = ops.score_topk(uw.to(dev), iw.to(dev), k, seen, u0=37, u1=U - 5, precise=True)
        assert torch.equal(idx2, ridx[37:U - 5]) and torch.equal(val2, rval[37:U - 5])


def test_score_topk_adversarial_orders(dev):
    """Scores increasing with the item id (every new item beats the threshold: the buffers compact at the maximum
    rate), decreasing, and all-equal (pure tie-break) rows."""
    from lgcnhs_b200 import ops

    U, M, k = 70, 4000, 20
    uw = torch.zeros(U, 64)
    iw = torch.zeros(M, 64)
    uw[:, 0] = 1.0
    uw[1::3, 0] = -1.0                                  # decreasing rows
    uw[2::3, 0] = 0.0                                   # all-equal rows
    iw[:, 0] = torch.arange(M, dtype=torch.float32) / M
    idx, val = ops.score_topk(uw.to(dev), iw.to(dev), k, precise=True)
    ref = uw @ iw.T
    rv, _ = torch.topk(ref, k)
    assert torch.equal(val.cpu(), rv)
    idx = idx.cpu()
    assert idx[0].tolist() == list(range(M - 1, M - 1 - k, -1))
    assert idx[1].tolist() == list(range(0, k)) or torch.equal(torch.gather(ref, 1, idx)[1], rv[1])
    assert idx[2].tolist() == list(range(M - 1, M - 1 - k, -1))      # ties -> larger index first


def test_score_topk_fused_hadamard_and_exclusion(dev):
    """F1: top-k of (G_score * F) with seen items dropped == reference getResourceMat + recommendForAllUser
    (model/SpreadLightGCN/model.py:151, recommend.py:34-46)."""
    from lgcnhs_b200 import ops

    U, M, k = 150, 2100, 30
    uw, iw, su, si = _rand_problem(U, M, 64, 5, 3000)
    g = torch.Generator().manual_seed(9)
    F = torch.rand(U, M, generator=g)
    F[F < 0.5] = 0.0
    seen = ops.seen_csr(su.to(dev), si.to(dev), U, M)
    idx, val = ops.score_topk(uw.to(dev), iw.to(dev), k, seen, exclude_seen=True, mul=F.to(dev), precise=True)
    # unfused device path: masked score, Hadamard, filtered top-k
    dense = ops.score_block(uw.to(dev), iw.to(dev), 0, U, seen) * F.to(dev)
    excl = ops.ExclusionMask.from_csr(seen, U, M)
    ridx, rval = ops.topk_rows(dense.contiguous(), k, excl)
    assert torch.equal(val, rval) and torch.equal(idx, ridx)
    # reference rule in float64 on the host
    G = (uw @ iw.T).double()
    G[su, si] = -1024.0
    Fn = G * F.double()
    Fn[su, si] = -np.inf
    rv, ri = torch.topk(Fn, k)
    assert_close(val, rv.float(), "fused Hadamard top-k scores")
    seen_set = set(zip(su.tolist(), si.tolist()))
    assert not any((u, int(i)) in seen_set for u in range(U) for i in idx[u].cpu())


@pytest.mark.parametrize("threads", [256, 512])
@pytest.mark.parametrize("k", [20, 50, 100])
def test_score_topk_large_user_tiles(dev, k, threads):
    """>= 128 * n_sms users selects the 128-user CTA tile (256 threads: 8 x 8 register tile, 512 threads: 4 x 8);
    k > 32 falls back to 64-user tiles."""
    from lgcnhs_b200 import ops
    from lgcnhs_b200._lib import lib

    U, M = 19100, 389
    uw, iw, su, si = _rand_problem(U, M, 64, 77, 40000)
    seen = ops.seen_csr(su.to(dev), si.to(dev), U, M)
    dense = ops.score_block(uw.to(dev), iw.to(dev), 0, U, seen)
    ridx, rval = ops.topk_rows(dense, k)
    default = 512
    lib().lgc_score_topk_config(threads)
    try:
        idx, val = ops.score_topk(uw.to(dev), iw.to(dev), k, seen, precise=True)
    finally:
        lib().lgc_score_topk_config(default)
    assert torch.equal(val, rval) and torch.equal(idx, ridx)


# ------------------------------------------------------------------------------------------
# the same contract on the tensor cores (lgc_score_topk_tc: tcgen05 kind::tf32, 3xTF32 split)
# ------------------------------------------------------------------------------------------
def _tc_check(val, idx, ref64, k, what, seen_mask=None):
    """Tolerance of the tensor-core path (north_star): score at rank within 1e-5 relative (+1e-7 of the largest ranked
    score), ids identical except at float near-ties, no excluded id."""
    from _parity import assert_topk_parity

    S = ref64.numpy().copy()
    if seen_mask is not None:
        S[seen_mask] = -np.inf
    rv, ri = torch.topk(torch.from_numpy(S), k)
    Sf = ref64.numpy()
    frac = assert_topk_parity(idx.cpu().numpy(), ri.numpy(), Sf, what, seen_mask=seen_mask, min_checked=0.6, tol_mult=1.0)
    got_v = val.cpu().double().numpy()
    want = np.take_along_axis(Sf, idx.cpu().numpy(), axis=1)
    scale = np.abs(want).max()
    assert (np.abs(got_v - want) <= 1e-5 * np.abs(want) + 1e-7 * scale).all(), \
        f"{what}: returned values off by {np.abs(got_v - want).max():.3e}"
    return frac


@pytest.mark.parametrize("U,M,dim,k", [(943, 1682, 64, 20), (300, 3706, 64, 10), (129, 700, 64, 32), (1000, 513, 32, 20),
                                       (6040, 3706, 64, 20), (2000, 26744, 64, 1)])
def test_score_topk_tensor_core_matches_fp64(dev, U, M, dim, k):
    """fill(-1024) semantics of recommend.py:86-114 on the tensor cores vs the float64 statement of the rule, and vs the
    fp32-FMA kernel (ids equal outside near-ties)."""
    from lgcnhs_b200 import ops

    uw, iw, su, si = _rand_problem(U, M, dim, U * 3 + M, 30 * U)
    seen = ops.seen_csr(su.to(dev), si.to(dev), U, M)
    idx, val = ops.score_topk(uw.to(dev), iw.to(dev), k, seen)                   # tensor-core path (U > 128, k <= 32)
    ref = uw.double() @ iw.double().T
    ref[su, si] = -1024.0
    _tc_check(val, idx, ref, k, f"tc score_topk {U}x{M}x{dim}")
    pidx, pval = ops.score_topk(uw.to(dev), iw.to(dev), k, seen, precise=True)
    assert (idx == pidx).float().mean() > 0.995
    assert torch.allclose(val, pval, rtol=1e-5, atol=1e-7 * float(pval.abs().max()))
    # user sub-range with an odd offset: rows are independent, so the block equals the slice bit for bit
    idx2, val2 = ops.score_topk(uw.to(dev), iw.to(dev), k, seen, u0=37, u1=U - 5)
    if U - 42 > 128:
        assert torch.equal(idx2, idx[37:U - 5]) and torch.equal(val2, val[37:U - 5])


def test_score_topk_tensor_core_seen_rule_and_ties(dev):
    """Seen pairs inside the top-k region: users with so few unseen items that -1024 entries must appear in the list
    (reference quirk: masked pairs stay in the ranking), exact ties (duplicate item rows -> larger id first), and the
    exclusion variant."""
    from lgcnhs_b200 import ops

    U, M, k = 300, 600, 20
    g = torch.Generator().manual_seed(4)
    uw = torch.empty(U, 64).normal_(std=0.1, generator=g)
    iw = torch.empty(M, 64).normal_(std=0.1, generator=g)
    iw[301] = iw[300]                                       # exact tie for every user
    dense_seen = torch.rand(U, M, generator=g) < 0.2
    dense_seen[:4, 5:] = True                               # four users have only items 0..4 unseen
    dense_seen[:4, :5] = False
    su, si = torch.nonzero(dense_seen, as_tuple=True)
    seen = ops.seen_csr(su.to(dev), si.to(dev), U, M)
    idx, val = ops.score_topk(uw.to(dev), iw.to(dev), k, seen)
    ref = uw.double() @ iw.double().T
    ref[su, si] = -1024.0
    rv, ri = torch.topk(ref, k)
    assert torch.allclose(val.cpu().double(), rv, rtol=1e-5, atol=1e-6)
    assert (val[:4, 5:] == -1024.0).all() and (val[:4, :5] > -1024.0).all()
    both = ((idx.cpu() == 300) | (idx.cpu() == 301)).sum(1) == 2
    pos300 = (idx.cpu() == 300).float().argmax(1)
    pos301 = (idx.cpu() == 301).float().argmax(1)
    assert (pos301[both] < pos300[both]).all() and both.any()       # equal scores: the larger id ranks first
    # exclusion variant + multiplier (fusion F1)
    F = torch.rand(U, M, generator=g)
    F[F < 0.5] = 0.0
    eidx, eval_ = ops.score_topk(uw.to(dev), iw.to(dev), k, seen, exclude_seen=True, mul=F.to(dev))
    Fn = ref.clone()
    Fn[su, si] = 0.0
    Fn = (uw.double() @ iw.double().T) * F.double()
    mask = dense_seen.numpy()
    _tc_check(eval_[4:], eidx[4:], Fn[4:], k, "tc fused Hadamard + exclusion", seen_mask=mask[4:])
    assert (eidx[:4, 5:] == -1).all()                       # fewer than k selectable items -> (-1, -inf) padding
    assert not dense_seen[torch.arange(U)[:, None].expand(-1, k)[eidx.cpu() >= 0], eidx.cpu()[eidx.cpu() >= 0]].any()


def test_score_topk_tensor_core_adversarial_orders(dev):
    """Scores increasing with the item id (every tile beats the threshold: maximum compaction rate), decreasing, and
    all-equal rows (pure tie-break), with values exactly representable so that the 3xTF32 split is exact."""
    from lgcnhs_b200 import ops

    U, M, k = 300, 4000, 20
    uw = torch.zeros(U, 64)
    iw = torch.zeros(M, 64)
    uw[:, 0] = 1.0
    uw[1::3, 0] = -1.0
    uw[2::3, 0] = 0.0
    iw[:, 0] = torch.arange(M, dtype=torch.float32) / 4096.0
    idx, val = ops.score_topk(uw.to(dev), iw.to(dev), k)
    ref = uw @ iw.T
    rv, _ = torch.topk(ref, k)
    assert torch.equal(val.cpu(), rv)
    idx = idx.cpu()
    assert idx[0].tolist() == list(range(M - 1, M - 1 - k, -1))
    assert idx[1].tolist() == list(range(0, k))
    assert idx[2].tolist() == list(range(M - 1, M - 1 - k, -1))
