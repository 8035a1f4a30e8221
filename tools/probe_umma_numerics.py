"""Diagnostics for the tcgen05 plane GEMM (run on the B200 box): where do errors sit, and how
does the tensor core round when it accumulates bf16 products into fp32?"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "light-graph-convolutional-recommendation-algorithm-based-on-hybrid-spreading_b200"))
import torch
from lgcnhs_b200 import ops

dev = torch.device("cuda:0")

def split(w, planes):
    out, r = [], w.clone()
    for _ in range(planes):
        h = r.to(torch.bfloat16); out.append(h); r = r - h.float()
    return torch.stack(out)

def case(M, N, K, planes, heavy=True):
    g = torch.Generator().manual_seed(M + N + K)
    ld = (K + 63) // 64 * 64
    A = torch.zeros(M, ld); A[:, :K] = (torch.rand(M, K, generator=g) < 0.2).float()
    W = torch.rand(N, K, generator=g)
    if heavy: W = W * torch.exp(torch.randn(N, K, generator=g) * 3)
    B = torch.zeros(planes, N, ld, dtype=torch.bfloat16); B[:, :, :K] = split(W, planes)
    Ad, Bd = A.to(torch.bfloat16).to(dev), B.to(dev)
    ref = A[:, :K].double() @ B[:, :, :K].double().sum(0).T
    C = ops.gemm_planes(0, Ad, Bd, M, N, K).cpu().double()
    Cs = ops.gemm_planes(0, Ad, Bd, M, N, K, simt=True).cpu().double()
    rel = ((C - ref).abs() / ref.abs().clamp_min(1e-30))
    bad = rel > 1e-5
    print(f"case M={M} N={N} K={K} planes={planes} heavy={heavy}: max rel {rel.max():.3e}, bad {int(bad.sum())}, "
          f"simt max rel {((Cs-ref).abs()/ref.abs().clamp_min(1e-30)).max():.3e}")
    if bad.any():
        r, c = bad.nonzero(as_tuple=True)
        print("   bad rows uniq", r.unique().numel(), "cols uniq", c.unique().numel(), "row range", int(r.min()), int(r.max()),
              "col range", int(c.min()), int(c.max()))
        print("   first bad:", [(int(a), int(b), float(C[a, b]), float(ref[a, b])) for a, b in list(zip(r, c))[:5]])
        # per-plane check: each plane alone
        for p in range(planes):
            Cp = ops.gemm_planes(0, Ad, Bd[p:p+1], M, N, K).cpu().double()
            rp = A[:, :K].double() @ B[p, :, :K].double().T
            e = (Cp - rp).abs()
            print(f"   plane {p}: max abs err {e.max():.3e} (max |ref| {rp.abs().max():.3e}), rel-to-rowmax {(e / rp.abs().max()).max():.3e}")

for args in [(333, 1111, 4100, 3, True), (333, 1111, 4100, 3, False), (333, 1111, 4100, 1, True), (128, 80, 4100, 3, True),
             (333, 1111, 2048, 3, True), (333, 1111, 8192, 3, False)]:
    case(*args)

# ---- accumulation rounding probes: one row of ones times B = [big, 1, 1, ...] ----
print("accumulate probes (A = ones(128,K), B row = [2^e, 1 x (K-1)]), result - 2^e :")
for K in (16, 64, 256):
    A = torch.ones(128, 64 * ((K + 63) // 64), dtype=torch.bfloat16); A[:, K:] = 0
    res = []
    for e in (8, 16, 20, 23, 24, 25, 26, 28, 30, 34):
        B = torch.zeros(1, 16, A.shape[1], dtype=torch.bfloat16)
        B[0, :, :K] = 1.0; B[0, :, 0] = 2.0 ** e
        C = ops.gemm_planes(0, A.to(dev), B.to(dev), 128, 16, K).cpu().double()
        res.append((e, float(C[0, 0] - 2.0 ** e)))
    print("  K=%d:" % K, res, " exact would be", K - 1)
# big in the LAST k position (arrives after the small ones were accumulated)
A = torch.ones(128, 256, dtype=torch.bfloat16)
for e in (24, 26, 30):
    B = torch.zeros(1, 16, 256, dtype=torch.bfloat16); B[0] = 1.0; B[0, :, 255] = 2.0 ** e
    C = ops.gemm_planes(0, A.to(dev), B.to(dev), 128, 16, 256).cpu().double()
    print("  big last, e=%d: result-2^e = %g (exact 255)" % (e, float(C[0, 0] - 2.0 ** e)))
