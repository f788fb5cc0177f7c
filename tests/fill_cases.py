"""Shared bodies of the fillComplete parity tests: the same cases run against the CPU execution model of the kernels
(tests/test_emul_fill.py) and against the GPU through the C ABI (tests/test_gpu_fill.py).  `fill` is a callable
(entry_ptr, gids, vals, gmin, gmax, dom_min, dom_n) -> dict(rowptr, colind (0-based), vals, colmap, rowmax, n_cols,
n_local_used, n_same, nnz).  Everything is compared BIT-EXACTLY with the CPU oracle's fillComplete."""
import numpy as np

from helpers import O, global_rows, oracle_problem, problem_rows


def check_rank(fill, A, pid, rows_rp_cols_vals, n_global, dom_info=None):
    rows, rp, cols, vals = rows_rp_cols_vals
    info = dom_info or A.row_map.info(pid)
    got = fill(rp, cols, vals, 1, n_global, info["minMyGID"] if info["numMyElements"] else 1, info["numMyElements"])
    orp, oci, ova = A.csr(pid)
    assert np.array_equal(got["colmap"], A.col_map.my_gids(pid)), (pid, got["colmap"][:12], A.col_map.my_gids(pid)[:12])
    assert np.array_equal(got["rowptr"] + 1, orp)
    assert np.array_equal(got["colind"] + 1, oci)
    assert np.array_equal(got["vals"], ova)  # duplicates are summed in insertion order: bit-exact
    imp = A.importer
    assert got["n_same"] == (imp.field(pid, "numSameIDs") if imp is not None else got["n_cols"])
    n = len(rp) - 1
    exp_max = np.array([oci[orp[r] - 1:orp[r + 1] - 1].max() - 1 if orp[r + 1] > orp[r] else -1 for r in range(n)], dtype=np.int32)
    if got.get("rowmax") is not None:
        assert np.array_equal(got["rowmax"], exp_max)
    return got


def synthetic_case(fill, kind, N, P):
    c, m, A = oracle_problem(kind, N, P)
    n = global_rows(kind, N)
    for pid in range(1, P + 1):
        info = m.info(pid)
        rows = problem_rows(kind, N, info["minMyGID"], info["maxMyGID"])
        got = check_rank(fill, A, pid, rows, n)
        if kind != "powerlaw":
            assert got["nnz"] == len(rows[2])  # no duplicates: the sorted arrays are adopted as the CSR (no pack pass)


def custom_problem(P, n, row_lengths_fn, col_fn, seed, domain_counts=None):
    """rows of arbitrary length with duplicate and unsorted columns, inserted into the oracle and returned per rank"""
    rng = np.random.default_rng(seed)
    c = O.OComm(P)
    m = O.OMap.uniform(c, n)
    dom = O.OMap.linear(c, n, domain_counts) if domain_counts is not None else m
    per_rank, counts = [], []
    for pid in range(1, P + 1):
        info = m.info(pid)
        rows = np.arange(info["minMyGID"], info["maxMyGID"] + 1, dtype=np.int64)
        lens = np.array([row_lengths_fn(int(g), rng) for g in rows], dtype=np.int64)
        rp = np.zeros(len(rows) + 1, dtype=np.int64)
        np.cumsum(lens, out=rp[1:])
        cols = np.concatenate([col_fn(int(g), int(l), rng) for g, l in zip(rows, lens)] + [np.zeros(0, dtype=np.int64)]).astype(np.int64)
        vals = rng.standard_normal(len(cols))
        per_rank.append((rows, rp, cols, vals))
        counts.append(lens)
    A = O.OMatrix(m, counts)
    for pid, (rows, rp, cols, vals) in enumerate(per_rank, start=1):
        if len(rows):
            A.insert_rows(pid, rows, rp, cols, vals)
    A.fill_complete(dom, m)
    return c, m, dom, A, per_rank


def duplicates_long_rows_case(fill):
    n, P = 2600, 2
    # columns drawn from a narrow window -> many duplicates; two rows longer than kWarpRowMax (block-per-row kernel),
    # empty rows, and local columns nobody references (numSameIDs < numMyElements -> permuted local part)
    def lens(g, rng):
        if g in (7, 1900):
            return 1500 if g == 7 else 2300
        return 0 if g % 11 == 0 else int(rng.integers(1, 40))

    def cols(g, l, rng):
        if l > 1000:
            return rng.integers(1, n + 1, size=l) // 3 * 3 + 1  # every column = 1 mod 3: duplicates + unused columns
        return (np.clip(g + rng.integers(-30, 31, size=l), 1, n) - 1) // 2 * 2 + 1  # odd columns only

    c, m, dom, A, per_rank = custom_problem(P, n, lens, cols, seed=3)
    for pid in range(1, P + 1):
        got = check_rank(fill, A, pid, per_rank[pid - 1], n)
        assert got["nnz"] < len(per_rank[pid - 1][2])  # duplicates were merged: pack path
        assert got["n_same"] < m.info(pid)["numMyElements"]


def domain_differs_case(fill):
    n, P = 90, 4
    dom_counts = [40, 0, 30, 20]  # rank 2 owns no domain ids
    c, m, dom, A, per_rank = custom_problem(P, n, lambda g, rng: int(rng.integers(1, 9)),
                                            lambda g, l, rng: rng.integers(1, n + 1, size=l), seed=11, domain_counts=dom_counts)
    for pid in range(1, P + 1):
        check_rank(fill, A, pid, per_rank[pid - 1], n, dom_info=dom.info(pid))


def empty_and_bad_gid_case(fill, error_type):
    got = fill([0], [], [], 1, 10, 1, 10)
    assert got["n_cols"] == 0 and got["nnz"] == 0 and got["rowptr"].tolist() == [0]
    got = fill([0, 0, 0], [], [], 1, 10, 1, 10)
    assert got["rowptr"].tolist() == [0, 0, 0]
    try:
        fill([0, 2], [3, 11], [1.0, 2.0], 1, 10, 1, 10)
    except error_type as e:
        assert "outside" in str(e)
    else:
        raise AssertionError("a column GID outside the global range must be rejected")


def overlapping_rows_problem(P, n_per, seed=5, width=3):
    """A matrix assembled on an OVERLAPPING row map (every rank holds its own n_per rows plus the first two rows of the
    next rank, with its own partial contributions to them -- finite-element style), domain = range = the uniform map:
    rangeMap != rowMap, so apply! takes the exporter branch.  Returns (comm, uniform map, row map,
    filled OMatrix, per-rank (rows, rowptr0, cols, vals), dense global operator)."""
    rng = np.random.default_rng(seed)
    N = P * n_per
    c = O.OComm(P)
    uni = O.OMap.uniform(c, N)

    def row_gids(p):
        own = list(range((p - 1) * n_per + 1, p * n_per + 1))
        nxt = [(p % P) * n_per + 1, (p % P) * n_per + 2] if P > 1 else []
        return own + nxt

    rowmap = O.OMap.from_gids(c, -1, [row_gids(p) for p in range(1, P + 1)])
    dense = np.zeros((N, N))
    per_rank, counts = [], []
    for p in range(1, P + 1):
        rows = np.array(row_gids(p), dtype=np.int64)
        lens = rng.integers(1, 2 * width + 2, size=len(rows))
        rp = np.zeros(len(rows) + 1, dtype=np.int64)
        np.cumsum(lens, out=rp[1:])
        # the diagonal first (every global column is referenced, as the serial column map requires), then scattered columns
        cols = np.concatenate([np.concatenate([[g], np.clip(g + rng.integers(-width * n_per // 2, width * n_per // 2 + 1, size=l - 1), 1, N)])
                               for g, l in zip(rows, lens)])
        vals = rng.standard_normal(len(cols))
        for r, g in enumerate(rows):
            for i in range(rp[r], rp[r + 1]):
                dense[g - 1, cols[i] - 1] += vals[i]
        per_rank.append((rows, rp, cols.astype(np.int64), vals))
        counts.append(lens)
    A = O.OMatrix(rowmap, counts)
    for p, (rows, rp, cols, vals) in enumerate(per_rank, start=1):
        A.insert_rows(p, rows, rp, cols, vals)
    A.fill_complete(uni, uni)
    return c, uni, rowmap, A, per_rank, dense
