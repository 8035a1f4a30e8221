"""GPU parity: (S1-S3, F1) tcgen05 GEMM, HybridS scaling, resource matrix, full recommend."""
import numpy as np
import pytest
import torch

from oracle import spread_oracle as S
from test_gpu_propagation import assert_close

pytestmark = pytest.mark.gpu


def _split_bf16(w, planes):
    out, r = [], w.clone()
    for _ in range(planes):
        h = r.to(torch.bfloat16)
        out.append(h)
        r = r - h.float()
    return torch.stack(out)


@pytest.mark.parametrize("M,N,K,planes", [(128, 256, 64, 1), (128, 80, 128, 3), (200, 300, 1000, 3), (1000, 517, 777, 2),
                                          (333, 1111, 4100, 3), (64, 64, 64, 1)])
def test_gemm_bf16_planes(dev, M, N, K, planes):
    from lgcnhs_b200 import ops

    g = torch.Generator().manual_seed(M + N + K)
    ld = (K + 63) // 64 * 64
    A = torch.zeros(M, ld)
    A[:, :K] = (torch.rand(M, K, generator=g) < 0.2).float()
    W = torch.rand(N, K, generator=g) * torch.exp(torch.randn(N, K, generator=g) * 3)
    B = torch.zeros(planes, N, ld, dtype=torch.bfloat16)
    B[:, :, :K] = _split_bf16(W, planes)
    rs = torch.rand(M, generator=g) + 0.5
    cs = torch.rand(N, generator=g) + 0.5
    ref = (A[:, :K].double() @ B[:, :, :K].double().sum(0).T) * rs.double()[:, None] * cs.double()[None, :] * 0.5
    Ad, Bd = A.to(torch.bfloat16).to(dev), B.to(dev)
    C = ops.gemm_planes(0, Ad, Bd, M, N, K, rs=rs.to(dev), cs=cs.to(dev), scale=0.5)
    assert_close(C, ref.float(), "tcgen05 bf16 planes vs fp64")
    Cs = ops.gemm_planes(0, Ad, Bd, M, N, K, rs=rs.to(dev), cs=cs.to(dev), scale=0.5, simt=True)
    assert_close(C, Cs, "tcgen05 vs on-device SIMT cross-check")


@pytest.mark.parametrize("M,N,K,planes", [(128, 64, 128, 4), (300, 200, 1000, 4), (517, 1000, 6040, 4), (100, 100, 300, 2),
                                          (90, 257, 130, 1)])
def test_gemm_u8_digits_exact(dev, M, N, K, planes):
    """Integer work: bit-exact against int64 arithmetic."""
    from lgcnhs_b200 import ops

    g = torch.Generator().manual_seed(M * 7 + N)
    ld = (K + 127) // 128 * 128
    A = torch.zeros(M, ld, dtype=torch.uint8)
    A[:, :K] = (torch.rand(M, K, generator=g) < 0.3).to(torch.uint8)
    B = torch.zeros(planes, N, ld, dtype=torch.uint8)
    B[:, :, :K] = torch.randint(0, 256, (planes, N, K), generator=g, dtype=torch.uint8)
    tot = torch.zeros(M, N, dtype=torch.int64)
    for p in range(planes):
        tot += (A[:, :K].long() @ B[p, :, :K].long().T) << (8 * p)
    scale = 2.0 ** -20
    ref = (tot.double() * scale).float()
    C = ops.gemm_planes(1, A.to(dev), B.to(dev), M, N, K, scale=scale)
    assert torch.equal(C.cpu(), ref)


@pytest.mark.parametrize("w_mode", ["u8x4", "u8x3", "bf16x3", "bf16x2"])
@pytest.mark.parametrize("name,lam", [("tiny", 0.3), ("small", 0.0), ("small", 1.0), ("ml-100k", 0.3), ("ml-100k", 0.85)])
def test_spreading_pipeline(dev, name, lam, w_mode):
    """G, W, F and the filtered top-k against the float64 reference restatement."""
    from lgcnhs_b200 import ops
    from lgcnhs_b200.synth import synth_shape

    d = synth_shape(name)
    tr, va, _ = d.split()
    sel = np.concatenate([tr, va])
    users, items = d.users[sel], d.items[sel]
    A = S.interaction_matrix(d.n_users, d.n_items, users, items)
    G = S.get_spreading_general_mat(A)
    W = S.hybrids(A, G, lam)
    F = S.get_resource(A, W)
    eng = ops.SpreadingEngine(d.n_users, d.n_items, torch.from_numpy(users).to(dev), torch.from_numpy(items).to(dev),
                              w_mode=w_mode)
    assert np.array_equal(eng.ku.cpu().numpy(), A.sum(1).astype(np.int32))
    assert np.array_equal(eng.ki.cpu().numpy(), A.sum(0).astype(np.int32))
    Gd = eng.general_w()
    assert_close(Gd, torch.from_numpy(G), "G = A^T K_u^-1 A")
    W32 = eng.scale(lam, want_w32=True)
    assert_close(W32, torch.from_numpy(W), "W = HybridS")
    Fd = eng.resource()
    assert_close(Fd, torch.from_numpy(F), "F = A W")
    k = 20
    idx, val = ops.topk_rows(Fd, k, eng.excl)
    fi, fv = S.recommend_fast(F, A, k)
    assert_close(val, torch.from_numpy(fv), "score at rank")
    gap_ok = np.ones_like(fi, dtype=bool)                # ids identical except at (near-)ties
    tol = 1e-5 * np.abs(fv) + 1e-7 * np.abs(F).max()
    gap_ok[:, 1:] &= (fv[:, :-1] - fv[:, 1:]) > 2 * tol[:, 1:]
    gap_ok[:, :-1] &= (fv[:, :-1] - fv[:, 1:]) > 2 * tol[:, :-1]
    # the k-th/(k+1)-th boundary can also be a tie: exclude the last rank
    gap_ok[:, -1] = False
    assert np.array_equal(idx.cpu().numpy()[gap_ok], fi[gap_ok])


def test_spreading_mass_conservation(dev):
    """HybridS(lambda=0) then A.W conserves mass per user: F.sum(axis=1) == k_u."""
    from lgcnhs_b200 import ops
    from lgcnhs_b200.synth import synth_shape

    d = synth_shape("ml-100k")
    eng = ops.SpreadingEngine(d.n_users, d.n_items, torch.from_numpy(d.users).to(dev), torch.from_numpy(d.items).to(dev))
    eng.general_w()
    eng.scale(0.0)   # lambda = 0: W[i,j] = G[i,j]/k_i, every row of W sums to 1
    F = eng.resource()
    assert torch.allclose(F.sum(1).cpu(), eng.ku.float().cpu(), rtol=1e-5)


def test_fusion_hadamard(dev):
    from lgcnhs_b200 import ops
    from lgcnhs_b200._lib import lib, check

    g = torch.Generator().manual_seed(5)
    F = torch.rand(100, 333, generator=g)
    G = torch.randn(100, 333, generator=g)
    Fd = F.to(dev)
    check(lib().hs_hadamard(Fd.data_ptr(), G.to(dev).data_ptr(), 100, 333, 333, 333, torch.cuda.current_stream().cuda_stream))
    assert torch.equal(Fd.cpu(), F * G)


@pytest.mark.parametrize("kind,planes,M,N,K", [(0, 3, 700, 500, 2000), (0, 2, 257, 300, 640), (0, 1, 512, 130, 200),
                                               (1, 4, 900, 333, 5000), (1, 2, 300, 300, 300), (1, 1, 256, 64, 128)])
def test_gemm_cta_pair_equals_single_cta(dev, kind, planes, M, N, K):
    """The cta_group::2 kernel (UMMA M=256, half of B per CTA) must reproduce the single-CTA kernel bit for bit."""
    from lgcnhs_b200 import ops
    from lgcnhs_b200._lib import check, lib

    g = torch.Generator().manual_seed(kind * 100 + planes)
    if kind == 0:
        ld = (K + 63) // 64 * 64
        A = torch.zeros(M, ld)
        A[:, :K] = (torch.rand(M, K, generator=g) < 0.2).float()
        B = torch.zeros(planes, N, ld, dtype=torch.bfloat16)
        B[:, :, :K] = _split_bf16(torch.rand(N, K, generator=g) * torch.exp(torch.randn(N, K, generator=g) * 2), planes)
        Ad, Bd = A.to(torch.bfloat16).to(dev), B.to(dev)
    else:
        ld = (K + 127) // 128 * 128
        A = torch.zeros(M, ld, dtype=torch.uint8)
        A[:, :K] = (torch.rand(M, K, generator=g) < 0.3).to(torch.uint8)
        B = torch.zeros(planes, N, ld, dtype=torch.uint8)
        B[:, :, :K] = torch.randint(0, 256, (planes, N, K), generator=g, dtype=torch.uint8)
        Ad, Bd = A.to(dev), B.to(dev)
    try:
        check(lib().hs_gemm_use_cta_pair(0))
        C1 = ops.gemm_planes(kind, Ad, Bd, M, N, K, scale=0.25).clone()
        check(lib().hs_gemm_use_cta_pair(1))
        C2 = ops.gemm_planes(kind, Ad, Bd, M, N, K, scale=0.25)
        torch.cuda.synchronize()
    finally:
        lib().hs_gemm_use_cta_pair(1)
    assert torch.equal(C1, C2)


@pytest.mark.parametrize("name", ["small", "ml-100k", "ml-1m"])
def test_symmetric_g_schedule_is_bit_identical(dev, name):
    """hs_gemm_planes_sym computes only the tiles that touch the upper triangle and mirrors them: the result must be
    bit-identical to the full computation (and to its own transpose), at sizes with ragged last tiles."""
    from lgcnhs_b200 import ops
    from lgcnhs_b200.synth import synth_shape

    d = synth_shape(name)
    eng = ops.SpreadingEngine(d.n_users, d.n_items, torch.from_numpy(d.users).to(dev), torch.from_numpy(d.items).to(dev))
    operands = eng.pack_g_operands()
    full = eng.general_w(operands=operands, symmetric=False).clone()
    out = torch.full((d.n_items, (d.n_items + 3) // 4 * 4), float("nan"), device=dev)[:, : d.n_items]
    sym = eng.general_w(operands=operands, symmetric=True, out=out)
    assert not torch.isnan(sym).any()                       # every entry written (directly or by the mirror store)
    assert torch.equal(sym, full) and torch.equal(sym, sym.T)
    A = S.interaction_matrix(d.n_users, d.n_items, d.users, d.items)
    if name != "ml-1m":
        assert_close(sym, torch.from_numpy(S.get_spreading_general_mat(A)), "symmetric-schedule G vs oracle")


@pytest.mark.parametrize("name,k,w_mode", [("small", 10, "u8x4"), ("small", 32, "u8x3"), ("ml-100k", 20, "u8x4"),
                                           ("ml-100k", 1, "u8x4"), ("ml-1m", 20, "u8x4"), ("ml-1m", 20, "u8x3")])
def test_fused_resource_topk_equals_unfused(dev, name, k, w_mode):
    """hs_resource_topk (top-k inside the F-GEMM epilogue, F never written) must return EXACTLY the lists and values of
    hs_gemm_planes + lgc_topk_rows: same fixed-point arithmetic, same total order (value desc, larger column first)."""
    from lgcnhs_b200 import ops
    from lgcnhs_b200.synth import synth_shape

    d = synth_shape(name)
    tr, va, _ = d.split()
    sel = np.r_[tr, va]
    eng = ops.SpreadingEngine(d.n_users, d.n_items, torch.from_numpy(d.users[sel]).to(dev),
                              torch.from_numpy(d.items[sel]).to(dev), w_mode=w_mode)
    assert eng.can_fuse_topk(k, d.n_users)
    for lam in (0.0, 0.3, 1.0):
        for filtered in (True, False):
            i0, v0 = eng.recommend(lam, k, filtered=filtered, fused=False)
            i1, v1 = eng.recommend(lam, k, filtered=filtered, fused=True)
            assert torch.equal(i0, i1), f"{name} lam={lam} filtered={filtered}: {(i0 != i1).sum().item()} ids differ"
            assert torch.equal(v0, v1)
    # a block of users with an offset into the exclusion mask (the multi-GPU user-block shard)
    u0, u1 = d.n_users // 3, d.n_users // 3 + max(129, d.n_users // 2)
    u1 = min(u1, d.n_users)
    if u1 - u0 > 128:
        ia, va_ = eng.recommend(0.3, k, user_range=(u0, u1), fused=False)
        ib, vb = eng.recommend(0.3, k, user_range=(u0, u1), fused=True)
        assert torch.equal(ia, ib) and torch.equal(va_, vb)
    # against the oracle (tie-aware) on the small shapes
    if name != "ml-1m":
        from _parity import assert_topk_parity
        A = S.interaction_matrix(d.n_users, d.n_items, d.users[sel], d.items[sel])
        F = S.get_resource(A, S.hybrids(A, S.get_spreading_general_mat(A), 0.3))
        ref_idx, _ = S.recommend_fast(F, A, k)
        got, _ = eng.recommend(0.3, k, fused=True)
        assert_topk_parity(got.cpu().numpy(), ref_idx, F, f"fused resource top-k vs oracle ({name})", seen_mask=A > 0,
                           min_checked=0.5)


def test_fused_resource_topk_few_selectable_items(dev):
    """Rows with fewer than k selectable items get (-1, -inf) padding, exactly like lgc_topk_rows."""
    from lgcnhs_b200 import ops

    g = np.random.default_rng(1)
    U, M = 200, 40
    dense = g.random((U, M)) < 0.6
    dense[:5] = True                                        # five users have seen everything
    dense[5, : M - 3] = True                                # one user has three items left
    dense[5, M - 3:] = False
    u, i = np.nonzero(dense)
    eng = ops.SpreadingEngine(U, M, torch.from_numpy(u).to(dev), torch.from_numpy(i).to(dev))
    i0, v0 = eng.recommend(0.5, 20, fused=False)
    i1, v1 = eng.recommend(0.5, 20, fused=True)
    assert torch.equal(i0, i1) and torch.equal(v0, v1)
    assert (i1[:5] == -1).all() and (i1[5, 3:] == -1).all() and (i1[5, :3] >= M - 3).all()
