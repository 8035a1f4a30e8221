"""Seeded synthetic interaction graphs of the shapes BASELINE.json names.

There is no network for MovieLens / Douban / Amazon-Book, so benchmarks and parity tests use
graphs of the same shape: Zipf-like item popularity (p_i ∝ rank^-0.8), log-normal user
activity, every user with >= min_per_user interactions (MovieLens' own filter), E distinct
(user, item) pairs, dense 0-based ids as after LabelEncoder
(/root/reference/processing/handleData.py:70-73), split 80/10/10 with
sklearn.model_selection.train_test_split(random_state=42) exactly as handleData.py:88-94.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

SHAPES = {
    # name: (users, items, interactions)
    "tiny": (96, 160, 2400),
    "small": (300, 500, 12000),
    "ml-100k": (943, 1682, 100_000),
    "ml-1m": (6040, 3706, 1_000_209),
    "douban": (640, 16_000, 64_000),          # assumed post-filter size, SURVEY.md §8
    "amazon-book": (52_643, 91_599, 2_984_108),
    "ml-20m": (138_493, 26_744, 20_000_263),
}


@dataclass
class Interactions:
    n_users: int
    n_items: int
    users: np.ndarray  # int64 [E], distinct pairs
    items: np.ndarray

    def split(self, seed: int = 42):
        """(train, val, test) index arrays, 80/10/10 (handleData.py:88-94)."""
        from sklearn.model_selection import train_test_split

        idx = np.arange(self.users.shape[0])
        train, rest = train_test_split(idx, test_size=0.2, random_state=seed)
        val, test = train_test_split(rest, test_size=0.5, random_state=seed)
        return train, val, test


def synth(n_users: int, n_items: int, n_edges: int, seed: int = 42, min_per_user: int = 20,
          uniform: bool = False) -> Interactions:
    rng = np.random.default_rng(seed)
    min_per_user = min(min_per_user, n_items // 2, max(1, n_edges // (2 * n_users)))
    if uniform:
        key = rng.choice(n_users * n_items, size=n_edges, replace=False)
        return Interactions(n_users, n_items, key // n_items, key % n_items)
    p = np.arange(1, n_items + 1, dtype=np.float64) ** -0.8
    p = p[rng.permutation(n_items)]  # popularity not correlated with item id
    p /= p.sum()
    q = rng.lognormal(0.0, 1.0, n_users)
    q /= q.sum()
    cp = np.cumsum(p)
    cq = np.cumsum(q)
    keys = np.empty(0, dtype=np.int64)
    # floor: min_per_user distinct items for every user
    need = np.full(n_users, min_per_user, dtype=np.int64)
    for _ in range(64):
        todo = np.nonzero(need > 0)[0]
        if todo.size == 0:
            break
        u = np.repeat(todo, need[todo] + 2)
        i = np.searchsorted(cp, rng.random(u.size)).clip(0, n_items - 1)
        keys = np.unique(np.concatenate([keys, u * n_items + i]))
        cnt = np.bincount(keys // n_items, minlength=n_users)
        # trim users that overshot the floor so the floor pass stays a floor
        need = np.maximum(min_per_user - cnt, 0)
    # fill up to n_edges with pairs drawn ∝ q_u * p_i, rejecting duplicates
    while keys.size < n_edges:
        m = int((n_edges - keys.size) * 1.3) + 1024
        u = np.searchsorted(cq, rng.random(m)).clip(0, n_users - 1)
        i = np.searchsorted(cp, rng.random(m)).clip(0, n_items - 1)
        keys = np.unique(np.concatenate([keys, u * n_items + i]))
    if keys.size > n_edges:
        # drop surplus pairs only from positions above each user's floor
        perm = rng.permutation(keys.size)
        ku = keys[perm] // n_items
        order = np.argsort(ku, kind="stable")
        su = ku[order]
        first = np.searchsorted(su, su, side="left")
        rank = np.arange(su.size) - first
        droppable = perm[order][rank >= min_per_user]
        drop = rng.choice(droppable, size=keys.size - n_edges, replace=False)
        keep = np.ones(keys.size, dtype=bool)
        keep[drop] = False
        keys = keys[keep]
    keys = keys[rng.permutation(keys.size)]  # rating-file order is not sorted
    return Interactions(n_users, n_items, keys // n_items, keys % n_items)


def synth_shape(name: str, seed: int = 42, **kw) -> Interactions:
    u, m, e = SHAPES[name]
    return synth(u, m, e, seed=seed, **kw)


def bipartite_adj(n_users: int, users: np.ndarray, items: np.ndarray) -> np.ndarray:
    """(2, 2E) symmetric 'adjacency' COO in the (U+M) node space, row-major sorted — the format
    utils/graph.py:12-35 produces and LightGCN.forward(edge_index) receives."""
    n_items_off = items.astype(np.int64) + n_users
    rows = np.concatenate([users.astype(np.int64), n_items_off])
    cols = np.concatenate([n_items_off, users.astype(np.int64)])
    n = int(max(rows.max(), cols.max())) + 1
    key = np.unique(rows * n + cols)
    return np.stack([key // n, key % n])
