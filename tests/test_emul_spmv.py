"""The localApply kernels (csrc/jp_csr_kernels.cuh: spmv_kernel, spmm_kernel, spmm_runs_kernel, spmm_dia_kernel, every
block kind, TMA and cp.async staging, with and without the halo buffer) executed on the CPU model of tests/cuda_emul and
compared with the CPU oracle's localApply (src/CSRMatrix.jl:1123-1162).  DIRECT blocks are bit-exact; the rest meets the
1e-12 contract of the north star.  GPU runs of the same kernels: tests/test_gpu_parity.py."""
import numpy as np
import pytest

import emul
from helpers import O, oracle_problem, synthetic


def _problem(kind, N, **kw):
    c, m, A = oracle_problem(kind, N, 1, **kw)
    rp, ci, va = A.csr(1)
    return rp - 1, ci - 1, va


def _oracle(rp0, ci0, va, x, y0, alpha, beta):
    return O.local_apply_raw(y0.copy(order="F"), rp0 + 1, ci0 + 1, va, x, O.NO_TRANS, alpha, beta)


def _bound(rp0, ci0, va, x, y0, alpha, beta):
    return 1e-12 * (abs(alpha) * _oracle(rp0, ci0, np.abs(va), np.abs(x), np.zeros_like(y0), 1.0, 0.0) + abs(beta) * np.abs(y0)) + 1e-300


def _vectors(n_cols, n_rows, k):
    return synthetic.vector_block(1, n_cols, k, seed=1), synthetic.vector_block(1, n_rows, k, seed=2)


@pytest.mark.parametrize("tma", [1, 0])
@pytest.mark.parametrize("kind,N", [("5pt", 20), ("7pt", 9), ("27pt", 6)])
def test_spmv_direct_blocks_are_bit_exact(kind, N, tma):
    rp, ci, va = _problem(kind, N)
    n = len(rp) - 1
    x, y0 = _vectors(n, n, 1)
    for alpha, beta in ((1.0, 0.0), (3.0, 0.5)):
        got, st = emul.local_apply(rp, ci, va, x, y0, alpha, beta, tma=tma)
        assert st["direct"] == st["blocks"] > 0
        assert np.array_equal(got, _oracle(rp, ci, va, x, y0, alpha, beta))


@pytest.mark.parametrize("cap,allow_direct", [(2048, 1), (256, 1), (2048, 0)])
def test_spmv_irregular_rows_all_block_kinds(cap, allow_direct):
    rp, ci, va = _problem("powerlaw", 4000, mean=16.0, max_len=1500)
    n = len(rp) - 1
    x, y0 = _vectors(n, n, 1)
    got, st = emul.local_apply(rp, ci, va, x, y0, 3.0, 0.5, cap=cap, allow_direct=allow_direct)
    assert st["subwarp"] > 0 and (cap != 256 or st["longrow"] > 0)
    ref = _oracle(rp, ci, va, x, y0, 3.0, 0.5)
    assert (np.abs(got - ref) <= _bound(rp, ci, va, x, y0, 3.0, 0.5)).all()


def test_spmv_with_halo_buffer():
    rp, ci, va = _problem("7pt", 8)
    n = len(rp) - 1
    x, y0 = _vectors(n, n, 1)
    nloc = n - 2 * 64  # the last two z-planes play the remote columns
    got, st = emul.local_apply(rp, ci, va, x, y0, 3.0, 0.5, n_local_cols=nloc, halo_split=True)
    assert np.array_equal(got, _oracle(rp, ci, va, x, y0, 3.0, 0.5))
    rp, ci, va = _problem("powerlaw", 1500)
    n = len(rp) - 1
    x, y0 = _vectors(n, n, 1)
    got, st = emul.local_apply(rp, ci, va, x, y0, 3.0, 0.5, n_local_cols=1000, halo_split=True)
    ref = _oracle(rp, ci, va, x, y0, 3.0, 0.5)
    assert (np.abs(got - ref) <= _bound(rp, ci, va, x, y0, 3.0, 0.5)).all()


# gather kernel, run-staged kernel
@pytest.mark.parametrize("variant", [0, 6])
@pytest.mark.parametrize("k", [2, 8, 11, 32])
def test_spmm_kernels(variant, k):
    rp, ci, va = _problem("27pt", 6)
    n = len(rp) - 1
    x, y0 = _vectors(n, n, k)
    for halo in ((False,) if variant == 6 else (False, True)):  # the runs kernel serves launches without halo columns
        got, st = emul.local_apply(rp, ci, va, x, y0, 3.0, 0.5, variant=variant, T=2, n_local_cols=n - 40 if halo else None,
                                   halo_split=halo, cap=3456 if variant == 6 else 2048)
        assert got is not None
        ref = _oracle(rp, ci, va, x, y0, 3.0, 0.5)
        assert (np.abs(got - ref) <= _bound(rp, ci, va, x, y0, 3.0, 0.5)).all()


def test_spmm_irregular_rows():
    rp, ci, va = _problem("powerlaw", 2500, mean=16.0, max_len=900)
    n = len(rp) - 1
    x, y0 = _vectors(n, n, 5)
    for cap in (2048, 256):
        got, st = emul.local_apply(rp, ci, va, x, y0, 3.0, 0.5, cap=cap, T=4)
        ref = _oracle(rp, ci, va, x, y0, 3.0, 0.5)
        assert (np.abs(got - ref) <= _bound(rp, ci, va, x, y0, 3.0, 0.5)).all()
    got, st = emul.local_apply(rp, ci, va, x, y0, 3.0, 0.5, variant=6, cap=3456)
    assert got is None  # random columns: the run tables are rejected and the gather kernel is used


@pytest.mark.parametrize("kind,N", [("5pt", 20), ("7pt", 8)])
def test_spmm_runs_kernel_other_stencils(kind, N):
    rp, ci, va = _problem(kind, N)
    n = len(rp) - 1
    x, y0 = _vectors(n, n, 5)
    got, st = emul.local_apply(rp, ci, va, x, y0, 3.0, 0.5, variant=6, cap=3456)
    ref = _oracle(rp, ci, va, x, y0, 3.0, 0.5)
    assert (np.abs(got - ref) <= _bound(rp, ci, va, x, y0, 3.0, 0.5)).all()


def test_spmm_runs_kernels_with_more_than_32_runs_per_block():
    """40 far-apart double diagonals: every row block references 40 separate column runs, so the tile copy walks the
    run list in two passes of 32 lanes"""
    n, ndiag = 16000, 40
    rows = np.arange(n)
    offs = (np.arange(ndiag) - ndiag // 2) * 391
    cols = (rows[:, None, None] + offs[None, :, None] + np.arange(2)[None, None, :]).reshape(n, -1)
    ok = (cols >= 0) & (cols < n)
    rp = np.zeros(n + 1, dtype=np.int32)
    np.cumsum(ok.sum(axis=1), out=rp[1:])
    ci = cols[ok].astype(np.int32)
    rng = np.random.default_rng(2)
    va = rng.standard_normal(len(ci))
    x, y0 = _vectors(n, n, 6)
    ref = _oracle(rp, ci, va, x, y0, 3.0, 0.5)
    got, st = emul.local_apply(rp, ci, va, x, y0, 3.0, 0.5, variant=6, cap=3456)
    assert got is not None and (np.abs(got - ref) <= _bound(rp, ci, va, x, y0, 3.0, 0.5)).all()


def test_random_csr_structures_every_block_kind():
    """property test: random row-length profiles (empty rows, rows exactly as long as the stage, rows longer than it),
    random stage sizes / rows per block, with and without the halo split, TMA and cp.async staging, k = 1 and k > 1 --
    the gather kernels against the oracle"""
    import os
    from hypothesis import given, settings, strategies as st

    @settings(max_examples=int(os.environ.get("JP_HYPOTHESIS_EXAMPLES", "40")), deadline=None, derandomize=True)
    @given(st.integers(0, 2 ** 31), st.integers(1, 400), st.sampled_from([256, 512, 2048]), st.sampled_from([1, 7, 32, 256]),
           st.sampled_from([1, 3]), st.booleans(), st.booleans(), st.sampled_from([1, 2, 8, 32]))
    def run(seed, n_rows, cap, max_rows, k, tma, halo, T):
        rng = np.random.default_rng(seed)
        n_cols = int(rng.integers(1, 600))
        profile = rng.integers(0, 4)
        if profile == 0:
            lens = rng.integers(0, 9, size=n_rows)
        elif profile == 1:
            lens = rng.integers(0, 70, size=n_rows)
        elif profile == 2:
            lens = np.where(rng.random(n_rows) < 0.03, rng.integers(cap - 2, cap + 40, size=n_rows), rng.integers(0, 12, size=n_rows))
        else:
            lens = np.where(rng.random(n_rows) < 0.1, cap, rng.integers(0, 40, size=n_rows))
        rp = np.zeros(n_rows + 1, dtype=np.int32)
        np.cumsum(lens, out=rp[1:])
        ci = rng.integers(0, n_cols, size=int(rp[-1])).astype(np.int32)
        va = rng.standard_normal(len(ci))
        x = rng.standard_normal((n_cols, k))
        y0 = rng.standard_normal((n_rows, k))
        alpha, beta = (3.0, 0.5) if seed % 3 else (1.0, 0.0)
        nloc = int(rng.integers(0, n_cols + 1)) if halo else None
        got, stats = emul.local_apply(rp, ci, va, x, y0, alpha, beta, n_local_cols=nloc, halo_split=halo, cap=cap, max_rows=max_rows, T=T,
                                      tma=int(tma), allow_direct=int(seed % 2))
        ref = _oracle(rp, ci, va, x, y0, alpha, beta)
        assert (np.abs(got - ref) <= _bound(rp, ci, va, x, y0, alpha, beta)).all(), (stats, np.abs(got - ref).max())

    run()


def test_random_banded_structures_run_staged_kernels():
    """property test: random banded matrices (random band width, fill density, row count; empty rows; several bands) through
    the run-staged SpMM kernel; structures the run-table builder turns away (not banded enough) must be turned away, never
    mis-computed"""
    import os
    from hypothesis import given, settings, strategies as st

    @settings(max_examples=int(os.environ.get("JP_HYPOTHESIS_EXAMPLES", "30")), deadline=None, derandomize=True)
    @given(st.integers(0, 2 ** 31), st.integers(1, 300), st.integers(1, 12), st.sampled_from([2, 5, 9]))
    def run(seed, half_rows, width, k):
        variant = 6
        rng = np.random.default_rng(seed)
        n = 2 * half_rows  # even: X's columns stay 16-byte aligned
        nbands = int(rng.integers(1, 4))
        centers = np.concatenate([[0], rng.integers(-n, n, size=nbands - 1)])
        density = rng.uniform(0.5, 1.0)
        rows_cols = []
        for r in range(n):
            cols = np.concatenate([np.arange(r + c - width, r + c + width + 1) for c in centers])
            cols = np.unique(cols[(cols >= 0) & (cols < n)])
            keep = rng.random(len(cols)) < density
            rows_cols.append(cols[keep] if r % 17 else cols[:0])  # every 17th row is empty
        rp = np.zeros(n + 1, dtype=np.int32)
        np.cumsum([len(c) for c in rows_cols], out=rp[1:])
        ci = np.concatenate(rows_cols + [np.zeros(0, dtype=np.int64)]).astype(np.int32)
        va = rng.standard_normal(len(ci))
        x = rng.standard_normal((n, k))
        y0 = rng.standard_normal((n, k))
        got, stats = emul.local_apply(rp, ci, va, x, y0, 3.0, 0.5, variant=variant, cap=3456)
        if got is None:
            return
        ref = _oracle(rp, ci, va, x, y0, 3.0, 0.5)
        assert (np.abs(got - ref) <= _bound(rp, ci, va, x, y0, 3.0, 0.5)).all(), (variant, np.abs(got - ref).max())

    run()


def test_reference_golden_2x2_through_the_kernels():
    """test/CSRMatrixTests.jl:120-174: A = [[2,3],[5,7]], X = 2s (2 vectors): A*X == [10 10; 24 24], and the 1-D
    Laplacian-like rows of test/CSRMatrixMPITests.jl:29-63 on one rank (alpha = 3, beta = 0.5, Y0 = diag(1, 2))"""
    rp, ci, va = np.array([0, 2, 4]), np.array([0, 1, 0, 1]), np.array([2.0, 3.0, 5.0, 7.0])
    x = np.full((2, 2), 2.0)
    for variant in (0,):
        got, _ = emul.local_apply(rp, ci, va, x, np.zeros((2, 2)), 1.0, 0.0, variant=variant)
        assert np.array_equal(got, np.array([[10.0, 10.0], [24.0, 24.0]]))
    got, _ = emul.local_apply(rp, ci, va, x[:, :1], np.zeros((2, 1)), 1.0, 0.0)
    assert np.array_equal(got, np.array([[10.0], [24.0]]))
    # rank 1 of the 4-rank golden case: rows {1: [2, -1], 2: [-1, 2, -1]} on the column map [1, 2, 3]
    rp, ci, va = np.array([0, 2, 5]), np.array([0, 1, 0, 1, 2]), np.array([2.0, -1.0, -1.0, 2.0, -1.0])
    xcol = np.full((3, 2), 2.0)
    y0 = np.diag([1.0, 2.0])
    got, _ = emul.local_apply(rp, ci, va, xcol, y0, 3.0, 0.5, n_local_cols=2, halo_split=True)
    assert np.array_equal(got, 0.5 * y0 + np.array([[6.0, 6.0], [0.0, 0.0]]))  # test/CSRMatrixMPITests.jl:52-63


@pytest.mark.parametrize("kind,N,k", [("27pt", 12, 8), ("27pt", 10, 5), ("7pt", 14, 4), ("7pt", 12, 9), ("5pt", 46, 2), ("5pt", 40, 7)])
def test_spmm_diagonal_kernel(kind, N, k):
    """spmm_dia_kernel (diagonal-major copy built by dia_mark_kernel / dia_fill_kernel, window plan of jp_dia_plan.h): the rows
    it serves against the oracle, bit-identical to the run-staged kernel (same ascending-column FMA chain); the rows near
    the first / last row are left alone (the library hands them to the gather kernel).  Sizes that are no multiple of the
    128-row block exercise the partial last block, k not a multiple of 4 the partial vector tile."""
    rp, ci, va = _problem(kind, N)
    n = len(rp) - 1
    assert n % 2 == 0  # the TMA copies need an even leading dimension (the library takes the gather kernel otherwise)
    x, y0 = _vectors(n, n, k)
    for alpha, beta, parity in ((1.0, 0.0, 0), (3.0, 0.5, 1)):
        got, lo, hi = emul.spmm_dia(rp, ci, va, x, y0, alpha, beta, parity=parity)
        assert got is not None and hi - lo > n // 3 and lo % 2 == parity, (lo, hi, n)
        ref = _oracle(rp, ci, va, x, y0, alpha, beta)
        err = np.abs(got[lo:hi] - ref[lo:hi])
        assert (err <= _bound(rp, ci, va, x, y0, alpha, beta)[lo:hi]).all()
        assert np.array_equal(got[:lo], y0[:lo]) and np.array_equal(got[hi:], y0[hi:])  # untouched
        runs, _ = emul.local_apply(rp, ci, va, x, y0, alpha, beta, variant=6, cap=3456)
        if runs is not None:
            assert np.array_equal(got[lo:hi], runs[lo:hi])


def test_spmm_diagonal_kernel_rejects_irregular_matrices():
    rp, ci, va = _problem("powerlaw", 3000, mean=6.0, max_len=20)
    n = len(rp) - 1
    x, y0 = _vectors(n, n, 4)
    assert emul.spmm_dia(rp, ci, va, x, y0)[0] is None


@pytest.mark.parametrize("slab_cols", [64, 700, 5000])
@pytest.mark.parametrize("kind,N,kw", [("powerlaw", 4000, dict(mean=16.0, max_len=1500)), ("powerlaw", 1500, dict(mean=3.0, max_len=40)), ("7pt", 9, {})])
def test_spmv_column_slab_kernel(kind, N, kw, slab_cols):
    """spmv_cb_kernel: nonzeros in slab-major order, segmented warp sums, carries for runs that cross a 32-nonzero pass
    (rows with hundreds of entries in one slab span many passes, warps and CTAs), against the oracle; deterministic"""
    rp, ci, va = _problem(kind, N, **kw)
    n = len(rp) - 1
    x, y0 = _vectors(n, n, 1)
    for alpha, beta in ((1.0, 0.0), (3.0, 0.5), (-2.0, 1.0)):
        got = emul.spmv_cb(rp, ci, va, x, y0, alpha, beta, slab_cols=slab_cols)
        ref = _oracle(rp, ci, va, x, y0, alpha, beta)
        assert (np.abs(got - ref) <= _bound(rp, ci, va, x, y0, alpha, beta)).all()
    emul.set_schedule("random", 5)
    again = emul.spmv_cb(rp, ci, va, x, y0, -2.0, 1.0, slab_cols=slab_cols)
    emul.set_schedule("forward")
    assert np.array_equal(got, again)


def test_random_csr_structures_column_slab_kernel():
    """property test: random row-length profiles (empty rows, rows with hundreds of entries in one slab, a single row holding
    everything), columns in random order inside a row, random slab widths -- spmv_cb_kernel + spmv_cb_carry_kernel and the
    device builders against the oracle"""
    import os
    from hypothesis import given, settings, strategies as st

    @settings(max_examples=int(os.environ.get("JP_HYPOTHESIS_EXAMPLES", "40")), deadline=None, derandomize=True)
    @given(st.integers(0, 2 ** 31), st.integers(1, 500), st.sampled_from([1, 3, 17, 64, 1000]))
    def run(seed, n_rows, slab_cols):
        rng = np.random.default_rng(seed)
        n_cols = int(rng.integers(1, 700))
        profile = rng.integers(0, 4)
        if profile == 0:
            lens = rng.integers(0, 9, size=n_rows)
        elif profile == 1:
            lens = rng.integers(0, 70, size=n_rows)
        elif profile == 2:
            lens = np.where(rng.random(n_rows) < 0.03, rng.integers(300, 2500, size=n_rows), rng.integers(0, 6, size=n_rows))
        else:
            lens = np.zeros(n_rows, dtype=np.int64)
            lens[int(rng.integers(0, n_rows))] = int(rng.integers(1, 4000))
        rp = np.zeros(n_rows + 1, dtype=np.int32)
        np.cumsum(lens, out=rp[1:])
        ci = rng.integers(0, n_cols, size=int(rp[-1])).astype(np.int32)  # unsorted, with repeats: the kernel must not care
        va = rng.standard_normal(len(ci))
        x = rng.standard_normal((n_cols, 1))
        y0 = rng.standard_normal((n_rows, 1))
        alpha, beta = ((3.0, 0.5), (1.0, 0.0), (-1.5, 1.0))[seed % 3]
        got = emul.spmv_cb(rp, ci, va, x, y0, alpha, beta, slab_cols=min(slab_cols, n_cols))
        ref = _oracle(rp, ci, va, x, y0, alpha, beta)
        assert (np.abs(got - ref) <= _bound(rp, ci, va, x, y0, alpha, beta)).all(), np.abs(got - ref).max()

    run()


def test_random_diagonal_structures_diagonal_kernel():
    """property test: random sets of diagonals (isolated ones, adjacent groups, far apart, up to 32), random fill, random
    sizes and both window parities -- the plan of jp_dia_plan.h, the builders and spmm_dia_kernel against the oracle on the
    rows the kernel serves; the other rows must be left alone"""
    import os
    from hypothesis import given, settings, strategies as st

    @settings(max_examples=int(os.environ.get("JP_HYPOTHESIS_EXAMPLES", "30")), deadline=None, derandomize=True)
    @given(st.integers(0, 2 ** 31), st.integers(300, 900), st.integers(1, 32), st.sampled_from([1, 4, 7]), st.sampled_from([0, 1]))
    def run(seed, half_rows, ndiag, k, parity):
        rng = np.random.default_rng(seed)
        n = 2 * half_rows
        span = int(rng.integers(2, max(3, n // 3)))
        offs = np.unique(np.concatenate([[0], rng.integers(-span, span + 1, size=ndiag - 1)]))
        fill = rng.uniform(0.6, 1.0)
        rows = np.arange(n)
        cols = rows[:, None] + offs[None, :]
        ok = (cols >= 0) & (cols < n) & (rng.random(cols.shape) < fill)
        ok[:, list(offs).index(0)] = True
        for d in range(len(offs)):  # every offset must occur at least once, or the plan sees fewer diagonals -- harmless, but keep it exact
            r_ok = np.nonzero((cols[:, d] >= 0) & (cols[:, d] < n))[0]
            ok[r_ok[0], d] = True
        rp = np.zeros(n + 1, dtype=np.int32)
        np.cumsum(ok.sum(axis=1), out=rp[1:])
        ci = cols[ok].astype(np.int32)
        va = rng.standard_normal(len(ci))
        x = rng.standard_normal((n, k))
        y0 = rng.standard_normal((n, k))
        got, lo, hi = emul.spmm_dia(rp, ci, va, x, y0, 3.0, 0.5, parity=parity)
        if got is None:
            return  # too small for a single block of 128 rows with in-bounds windows
        assert lo % 2 == parity and 0 <= lo < hi <= n
        ref = _oracle(rp, ci, va, x, y0, 3.0, 0.5)
        assert (np.abs(got[lo:hi] - ref[lo:hi]) <= _bound(rp, ci, va, x, y0, 3.0, 0.5)[lo:hi]).all()
        assert np.array_equal(got[:lo], y0[:lo]) and np.array_equal(got[hi:], y0[hi:])

    run()
