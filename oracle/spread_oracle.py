"""ORACLE — TEST INFRASTRUCTURE ONLY.  Never import this from the product path.

CPU restatement (NumPy float64, like the reference) of the hybrid-spreading hot path
(S0-S5, F1).  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs may import it.

PARITY STATUS: the reference has no golden vectors of its own, but this half
of the reference is importable in the build container; oracle/make_golden.py runs the
reference's real model/SpreadMethod/model.py + recommend.py loop on seeded inputs, checks this
restatement against them bit for bit (same NumPy/BLAS), and commits the outputs as fixtures
under tests/golden/ — "pinned against outputs of the reference itself run here".

Each function cites the reference file:line it follows.
"""
from __future__ import annotations

import numpy as np


def interaction_matrix(user_num: int, item_num: int, users, items) -> np.ndarray:
    """utils/trans.py:13-29 (getInteractionMatrixByDataframe): A[u, i] = 1, float64."""
    A = np.zeros((user_num, item_num))
    A[np.asarray(users, dtype=np.int64), np.asarray(items, dtype=np.int64)] = 1
    return A


def get_spreading_general_mat(A: np.ndarray) -> np.ndarray:
    """model/SpreadMethod/model.py:14-27."""
    user_degrees = np.sum(A, axis=1)
    user_degrees[user_degrees == 0] = 1
    return np.dot(A.T / user_degrees, A)


def probs(A: np.ndarray, general_W: np.ndarray) -> np.ndarray:
    """model/SpreadMethod/model.py:30-43 (defined, never called by the reference)."""
    d = np.sum(A, axis=0)
    d[d == 0] = 1
    return general_W / d[np.newaxis, :]


def heats(A: np.ndarray, general_W: np.ndarray) -> np.ndarray:
    """model/SpreadMethod/model.py:46-60 (defined, never called by the reference)."""
    d = np.sum(A, axis=0)
    d[d == 0] = 1
    return general_W / d[:, np.newaxis]


def hybrids(A: np.ndarray, general_W: np.ndarray, Lambda: float) -> np.ndarray:
    """model/SpreadMethod/model.py:63-85."""
    item_degrees = np.sum(A, axis=0)
    degree_alpha = np.power(item_degrees, 1 - Lambda)
    degree_beta = np.power(item_degrees, Lambda)
    den = degree_alpha[:, np.newaxis] * degree_beta[np.newaxis, :]
    den[den == 0] = 1
    return general_W / den


def get_resource(A: np.ndarray, W: np.ndarray) -> np.ndarray:
    """model/SpreadMethod/model.py:88-99."""
    return np.dot(A, W)


def recommend_loop(F_new: np.ndarray, seen: dict, k: int, unfiltered: bool = False) -> dict:
    """model/SpreadMethod/recommend.py:35-50, the literal per-user loop (small cases only —
    32.8 ms/user at ML-1M shape).  `seen[uid]` = list of train+val items."""
    out = {}
    for uid in range(F_new.shape[0]):
        sorted_items = np.argsort(F_new[uid])[::-1]
        interacted = seen.get(uid, [])
        filtered = [iid for iid in sorted_items if iid not in interacted]
        out[uid] = filtered[:k]
        if unfiltered:                      # movielens + ProbS override, :49-50
            out[uid] = sorted_items[:k]
    return out


def recommend_fast(F_new: np.ndarray, A: np.ndarray, k: int, unfiltered: bool = False):
    """Same selection as recommend_loop, vectorised: rank by (value desc, index desc) — what
    np.argsort(row)[::-1] yields whenever the sort is stable on ties — after dropping the seen
    items.  Returns (idx (U,k) int64, values (U,k) float64)."""
    U, M = F_new.shape
    S = F_new.astype(np.float64, copy=True)
    if not unfiltered:
        S[A > 0] = -np.inf
    order = np.lexsort((np.broadcast_to(np.arange(M), (U, M)), S), axis=1)[:, ::-1][:, :k]
    return order.astype(np.int64), np.take_along_axis(F_new, order, axis=1)


def fused_resource(G_score: np.ndarray, F: np.ndarray) -> np.ndarray:
    """model/SpreadLightGCN/model.py:151 (== SpreadLightGCNOpti/model.py:241): F_new = G * F,
    G fp32 with -1024 at seen pairs, F float64."""
    return G_score * F
