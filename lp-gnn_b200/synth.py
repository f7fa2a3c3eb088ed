"""Seeded synthetic LPs of the shapes named in BASELINE.json (generator spec: SURVEY.md 8d).

A raw LP ``min c'x, b_l <= Ax <= b_u, l <= x <= u`` is drawn, then pushed through the same
scaling + feature pipeline the reference applies to real LPs (``features.scale_lp`` /
``features.node_features`` mirror dataset.py:23-96), so the result has the real processed-file
layout (dataset.py:213-217): ``row, col, A_data, c_feas[m,8], v_feas[n,8], y_s, y_t``.

Sparsity: variable j gets ``1 + Poisson(z/n - 1)`` nonzeros.  ``structure='staircase'`` places
them near the diagonal band ``floor(j*m/n) + U{-w..w}`` (MIRP-like time-staircase), ``'uniform'``
places them uniformly at random -- the locality-free worst case for the gather.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np
import scipy.sparse as sp

from .features import node_features, scale_lp

# (m, n, nnz, hids, depth) of BASELINE.json configs; C5 is a population, see lp_population().
CONFIGS = {
    "C1": dict(m=1_000, n=2_000, nnz=10_000, hids=64, depth=2, seed=1235),
    "C2": dict(m=50_000, n=100_000, nnz=500_000, hids=1024, depth=3, seed=1236),
    "C3": dict(m=50_000, n=100_000, nnz=500_000, hids=1024, depth=3, seed=1237),
    "C4": dict(m=1_000_000, n=2_000_000, nnz=10_000_000, hids=1024, depth=3, seed=1238),
}


@dataclass
class ProcessedLP:
    """The arrays of one processed ``.pk`` file (dataset.py:213-217) plus sizes."""
    row: np.ndarray      # int64 [z]  COO of the scaled A, row-major sorted
    col: np.ndarray      # int64 [z]
    a_data: np.ndarray   # float64 [z], |a| <= 1
    c_feas: np.ndarray   # float32 [m,8]
    v_feas: np.ndarray   # float32 [n,8]
    y_s: np.ndarray      # int64 [m] in {0,1,2}
    y_t: np.ndarray      # int64 [n]
    m: int
    n: int

    @property
    def nnz(self) -> int:
        return int(self.row.shape[0])


def raw_lp(m: int, n: int, nnz: int, seed: int, structure: str = "staircase", band: int = 64):
    """Returns ``(c, b_l, A_csr, b_u, l, u)`` float64."""
    rng = np.random.default_rng(seed)
    lam = max(nnz / n - 1.0, 0.0)
    k = 1 + rng.poisson(lam, size=n)
    cols = np.repeat(np.arange(n, dtype=np.int64), k)
    z = cols.shape[0]
    if structure == "staircase":
        centre = (cols * m) // n
        rows = (centre + rng.integers(-band, band + 1, size=z)) % m
    elif structure == "uniform":
        rows = rng.integers(0, m, size=z)
    else:
        raise ValueError(f"unknown structure {structure!r}")
    vals = np.where(rng.random(z) < 0.7, rng.choice([-1.0, 1.0], size=z), rng.uniform(-10, 10, size=z))
    vals[vals == 0.0] = 1.0
    # merge duplicates: keep the first value drawn for a (row, col) pair
    key = rows * n + cols
    _, first = np.unique(key, return_index=True)
    A = sp.csr_matrix((vals[first], (rows[first], cols[first])), shape=(m, n))
    A.sort_indices()

    rhs = rng.normal(0.0, 5.0, size=m)
    kind = rng.random(m)
    b_l = np.where(kind < 0.45, -np.inf, rhs)           # 45% "<=", 45% ">=", 10% "="
    b_u = np.where((kind >= 0.45) & (kind < 0.9), np.inf, rhs)
    l = np.zeros(n)
    u = np.where(rng.random(n) < 0.7, np.inf, rng.uniform(1.0, 10.0, size=n))
    c = rng.normal(0.0, 1.0, size=n)
    return c, b_l, A, b_u, l, u


def consistent_labels(c_feas: np.ndarray, v_feas: np.ndarray, rng) -> tuple[np.ndarray, np.ndarray]:
    """Random basis statuses that never contradict the +-inf tags (the invariant asserted at
    dataset.py:203-207): status 0 (at lower) impossible where the lower tag != 0, status 2
    (at upper) impossible where the upper tag != 0."""
    def draw(feas):
        y = rng.integers(0, 3, size=feas.shape[0])
        lo_inf, up_inf = feas[:, 5] != 0, feas[:, 7] != 0
        y = np.where((y == 0) & lo_inf, 1, y)
        y = np.where((y == 2) & up_inf, 1, y)
        return y.astype(np.int64)
    return draw(c_feas), draw(v_feas)


def processed_lp(m: int, n: int, nnz: int, seed: int, structure: str = "staircase") -> ProcessedLP:
    c, b_l, A, b_u, l, u = raw_lp(m, n, nnz, seed, structure)
    c, b_l, A, b_u, l, u = scale_lp(c, b_l, A, b_u, l, u)
    v_feas, c_feas = node_features(c, b_l, A, b_u, l, u)
    v_feas = v_feas.astype(np.float32)
    c_feas = c_feas.astype(np.float32)
    coo = A.tocoo()
    y_s, y_t = consistent_labels(c_feas, v_feas, np.random.default_rng(seed + 7919))
    return ProcessedLP(coo.row.astype(np.int64), coo.col.astype(np.int64), coo.data.astype(np.float64),
                       c_feas, v_feas, y_s, y_t, m, n)


def config_lp(name: str, structure: str = "staircase") -> ProcessedLP:
    cfg = CONFIGS[name]
    return processed_lp(cfg["m"], cfg["n"], cfg["nnz"], cfg["seed"], structure)


def lp_population(count: int, seed: int = 1239, m_lo: int = 100, m_hi: int = 20_000):
    """C5: sizes of ``count`` LPs, m log-uniform in [m_lo, m_hi], n = 2m, z = 5n, shuffled.
    Returns a list of ``(m, n, nnz, seed)``; LPs are materialised lazily by the caller."""
    rng = np.random.default_rng(seed)
    ms = np.exp(rng.uniform(np.log(m_lo), np.log(m_hi), size=count)).astype(np.int64)
    rng.shuffle(ms)
    return [(int(mm), int(2 * mm), int(10 * mm), int(seed * 100_003 + i)) for i, mm in enumerate(ms)]
