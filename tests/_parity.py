"""TEST INFRASTRUCTURE — shared parity checks (tolerances of SURVEY.md §8d / BASELINE.json north_star).

    floating point: |x - ref| <= 1e-5 |ref| + 1e-7 max|ref|              (embeddings, W, F)
    top-k lists   : score at every rank equal within that tolerance, ids identical except where the reference's
                    own neighbouring scores are closer than the tolerance (float ties), no excluded id returned
    metrics       : equal after the reference's round(., 5), +-1e-5
"""
import numpy as np

RTOL, ATOL_REL = 1e-5, 1e-7


def tol_of(ref: np.ndarray, scale: float) -> np.ndarray:
    return RTOL * np.abs(ref) + ATOL_REL * scale


def assert_topk_parity(got_idx, ref_idx, ref_scores, what="", seen_mask=None, min_checked=0.5, tol_mult=2.0):
    """got_idx / ref_idx: (U, k) item ids; ref_scores: (U, M) the REFERENCE's score matrix (fp64 or fp32).

    1. score at rank: ref_scores[u, got[u, r]] == ref_scores[u, ref[u, r]] within tolerance, for every rank;
    2. ids: identical wherever the reference's score at that rank is separated from both neighbours (and from the
       first item left out, i.e. rank k) by more than the tolerance;
    3. no id of `seen_mask` (bool (U, M), True = must not be recommended) is returned.
    Returns the fraction of (user, rank) positions whose id was compared."""
    got = np.asarray(got_idx, dtype=np.int64)
    ref = np.asarray(ref_idx, dtype=np.int64)
    S = np.asarray(ref_scores)
    assert got.shape == ref.shape, f"{what}: list shape {got.shape} vs {ref.shape}"
    U, k = ref.shape
    assert got.min() >= 0 and got.max() < S.shape[1], f"{what}: id out of range"
    fv = np.take_along_axis(S, ref, axis=1).astype(np.float64)
    gv = np.take_along_axis(S, got, axis=1).astype(np.float64)
    # absolute part of the tolerance: relative to the largest RANKED score, not to max|S| — the score matrices of the
    # LightGCN family hold the -1024 fill (and -1024 * F in the fusion), which would inflate it 1000-fold
    scale = float(np.abs(fv[np.isfinite(fv)]).max())
    tol = tol_mult * tol_of(fv, scale)
    bad = np.abs(gv - fv) > tol
    assert not bad.any(), f"{what}: score at rank differs at {int(bad.sum())} positions, max {np.abs(gv - fv).max():.3e}"
    # the best score NOT in the reference list (rank k), to know whether the last entry is tied with an outsider
    Sm = S.astype(np.float64, copy=True)
    if seen_mask is not None:
        Sm[seen_mask] = -np.inf
    np.put_along_axis(Sm, ref, -np.inf, axis=1)
    nxt = Sm.max(axis=1)
    ext = np.concatenate([np.full((U, 1), np.inf), fv, nxt[:, None]], axis=1)       # (U, k+2)
    gap_up = ext[:, :-2] - ext[:, 1:-1]
    gap_dn = ext[:, 1:-1] - ext[:, 2:]
    sure = (gap_up > 2 * tol) & (gap_dn > 2 * tol)
    assert np.array_equal(got[sure], ref[sure]), \
        f"{what}: {int((got[sure] != ref[sure]).sum())} ids differ outside float ties"
    for u in range(U):
        assert len(set(got[u].tolist())) == k, f"{what}: duplicate id in the list of user {u}"
    if seen_mask is not None:
        assert not np.take_along_axis(seen_mask, got, axis=1).any(), f"{what}: an excluded item was recommended"
    frac = float(sure.mean())
    assert frac >= min_checked, f"{what}: only {frac:.2%} of the positions are tie-free - the check is vacuous"
    return frac


def assert_close_np(x, ref, what="", extra=None):
    x = np.asarray(x, dtype=np.float64)
    ref = np.asarray(ref, dtype=np.float64)
    tol = tol_of(ref, float(np.abs(ref).max()))
    if extra is not None:
        tol = tol + extra
    bad = np.abs(x - ref) > tol
    assert not bad.any(), f"{what}: {int(bad.sum())} / {bad.size} out of tolerance, max err {np.abs(x - ref).max():.3e}"
