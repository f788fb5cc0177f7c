"""Synthetic, partition-independent inputs for the benchmarks and the parity tests.

Everything is a pure function of the global row id, so any rank (and the CPU oracle) can generate exactly its
own rows.  Grid row id g = i + N*(j + N*k) (0-based; GID = g+1).  Stencils: diagonal 4 / 6 / 26, off-diagonals
-1, zero-Dirichlet truncation (entries outside the grid are omitted); entries are listed in ascending GID order.
Vectors: u(g, seed) = (splitmix64(g xor seed*0x9E3779B97F4A7C15) >> 11) * 2^-53 * 2 - 1 in [-1, 1).
"""
from __future__ import annotations

import numpy as np

_GOLD = np.uint64(0x9E3779B97F4A7C15)


def splitmix64(x: np.ndarray) -> np.ndarray:
    with np.errstate(over="ignore"):
        z = (np.asarray(x, dtype=np.uint64) + _GOLD)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return z ^ (z >> np.uint64(31))


def uniform_pm1(g: np.ndarray, seed: int) -> np.ndarray:
    """u(g, seed) in [-1, 1) for 0-based global ids g"""
    with np.errstate(over="ignore"):
        key = np.asarray(g, dtype=np.uint64) ^ (np.uint64(seed) * _GOLD)
    return (splitmix64(key) >> np.uint64(11)).astype(np.float64) * (2.0 ** -53) * 2.0 - 1.0


def vector_block(gid_first: int, n: int, k: int, seed: int) -> np.ndarray:
    """column-major n x k block of the global multivector: vector j uses seed + 16*j"""
    g = np.arange(gid_first - 1, gid_first - 1 + n, dtype=np.uint64)
    out = np.empty((n, k), dtype=np.float64, order="F")
    for j in range(k):
        out[:, j] = uniform_pm1(g, seed + 16 * j)
    return out


STENCILS = {"5pt": (2, 4.0), "7pt": (3, 6.0), "27pt": (3, 26.0)}


def _offsets(kind: str):
    dim, _ = STENCILS[kind]
    if kind == "5pt":
        offs = [(0, -1), (-1, 0), (0, 0), (1, 0), (0, 1)]  # (di, dj), ascending di + N*dj
        return [(di, dj, 0) for di, dj in offs]
    if kind == "7pt":
        return [(0, 0, -1), (0, -1, 0), (-1, 0, 0), (0, 0, 0), (1, 0, 0), (0, 1, 0), (0, 0, 1)]
    return [(di, dj, dk) for dk in (-1, 0, 1) for dj in (-1, 0, 1) for di in (-1, 0, 1)]


def stencil_nnz(kind: str, N: int) -> int:
    if kind == "5pt":
        return 5 * N * N - 4 * N
    if kind == "7pt":
        return 7 * N ** 3 - 6 * N * N
    return (3 * N - 2) ** 3


def stencil_rows(kind: str, N: int, gid_first: int, gid_last: int):
    """rows gid_first..gid_last (1-based, inclusive) of the N^d stencil matrix.
    Returns (row GIDs, rowptr0, column GIDs, values) with per-row entries in ascending GID order."""
    dim, diag = STENCILS[kind]
    g = np.arange(gid_first - 1, gid_last, dtype=np.int64)
    i = g % N
    j = (g // N) % N
    k = g // (N * N) if dim == 3 else np.zeros_like(g)
    offs = _offsets(kind)
    cols = np.empty((len(g), len(offs)), dtype=np.int64)
    valid = np.empty((len(g), len(offs)), dtype=bool)
    vals = np.empty((len(g), len(offs)), dtype=np.float64)
    for t, (di, dj, dk) in enumerate(offs):
        ok = (i + di >= 0) & (i + di < N) & (j + dj >= 0) & (j + dj < N)
        if dim == 3:
            ok &= (k + dk >= 0) & (k + dk < N)
        valid[:, t] = ok
        cols[:, t] = g + di + N * dj + N * N * dk + 1
        vals[:, t] = diag if (di, dj, dk) == (0, 0, 0) else -1.0
    rowptr0 = np.zeros(len(g) + 1, dtype=np.int64)
    np.cumsum(valid.sum(axis=1), out=rowptr0[1:])
    return g + 1, rowptr0, cols[valid], vals[valid]


def powerlaw_row_lengths(r: np.ndarray, n: int, mean: float = 16.0, max_len: int = 10000, seed: int = 5) -> np.ndarray:
    """Pareto(a=1.5) row lengths clamp(floor(x_m * u^(-1/a)), 1, max_len) with x_m = mean/3 (the mean of the
    unclamped law is a*x_m/(a-1)); r are 0-based global row ids"""
    u = (uniform_pm1(r, seed) + 1.0) * 0.5
    u = np.maximum(u, 2.0 ** -53)
    xm = mean / 3.0
    ell = np.floor(xm * u ** (-1.0 / 1.5))
    return np.clip(ell, 1, min(max_len, n)).astype(np.int64)


def powerlaw_rows(n: int, gid_first: int, gid_last: int, mean: float = 16.0, max_len: int = 10000, seed: int = 3):
    """rows of the random irregular matrix: row r has its diagonal plus hashed columns
    c = 1 + splitmix64(r*2^32 + t) mod n (duplicates allowed: fillComplete merges them), values u(r*2^32+t, seed)."""
    r = np.arange(gid_first - 1, gid_last, dtype=np.int64)
    ell = powerlaw_row_lengths(r, n, mean, max_len)
    rowptr0 = np.zeros(len(r) + 1, dtype=np.int64)
    np.cumsum(ell, out=rowptr0[1:])
    total = int(rowptr0[-1])
    rowrep = np.repeat(r, ell)
    t = np.arange(total, dtype=np.int64) - np.repeat(rowptr0[:-1], ell)
    with np.errstate(over="ignore"):
        key = (rowrep.astype(np.uint64) << np.uint64(32)) + t.astype(np.uint64)
    cols = (splitmix64(key) % np.uint64(n)).astype(np.int64) + 1
    cols[rowptr0[:-1]] = r + 1  # the diagonal is always present (t == 0 slot)
    vals = uniform_pm1(key, seed)
    return r + 1, rowptr0, cols, vals


def build_matrix(jp, row_map, kind: str, N: int, chunk_rows: int = 1 << 21, fill: bool = True, **kw):
    """CSRMatrix of a synthetic problem on `row_map` (rows = this rank's GIDs of a uniform linear map), filled
    through the public insertGlobalValues path and fillComplete'd.  kind in {5pt, 7pt, 27pt, powerlaw}."""
    first, last = row_map.minMyGID(), row_map.maxMyGID()
    n_my = row_map.numMyElements()
    if kind == "powerlaw":
        counts = powerlaw_row_lengths(np.arange(first - 1, last, dtype=np.int64), N, kw.get("mean", 16.0), kw.get("max_len", 10000))
    else:
        counts = len(_offsets(kind))
    A = jp.CSRMatrix(row_map, counts, jp.STATIC_PROFILE)
    lo = first
    while lo <= last and n_my > 0:
        hi = min(last, lo + chunk_rows - 1)
        if kind == "powerlaw":
            rows, rp, cols, vals = powerlaw_rows(N, lo, hi, kw.get("mean", 16.0), kw.get("max_len", 10000))
        else:
            rows, rp, cols, vals = stencil_rows(kind, N, lo, hi)
        A.insertGlobalValuesBatch(rows, rp, cols, vals)
        lo = hi + 1
    if fill:
        A.fillComplete(row_map, row_map)
    return A
