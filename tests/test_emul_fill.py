"""Device fillComplete (csrc/jp_fill_kernels.cuh + jp_fill_pipeline.h) executed on the CPU model of tests/cuda_emul and
compared BIT-EXACTLY with the CPU oracle's fillComplete (column map, local indices, merged values, row pointers,
numSameIDs) for 1-4 simulated ranks.  The same cases run on the GPU in tests/test_gpu_fill.py."""
import numpy as np
import pytest

import emul
import fill_cases


@pytest.mark.parametrize("kind,N", [("5pt", 9), ("7pt", 5), ("27pt", 4), ("powerlaw", 700)])
@pytest.mark.parametrize("P", [1, 2, 3])
def test_synthetic_problems_match_oracle(kind, N, P):
    fill_cases.synthetic_case(emul.fill_complete, kind, N, P)


def test_duplicates_long_rows_and_unused_local_columns():
    fill_cases.duplicates_long_rows_case(emul.fill_complete)


def test_domain_map_differs_from_row_map_and_empty_rank():
    fill_cases.domain_differs_case(emul.fill_complete)


def test_empty_matrix_and_bad_gid():
    fill_cases.empty_and_bad_gid_case(emul.fill_complete, emul.EmulError)


def test_random_small_matrices_against_the_host_mirror():
    """property test: random row structures (empty rows, duplicates, unsorted and out-of-domain-but-valid columns, any
    split of the global ids into [below | domain | above]) -- the kernels' column map, local indices and merged values
    equal the host mirror's numpy fillComplete pieces (CSRMatrix._sortAndMerge and the documented column-map order)"""
    from hypothesis import given, settings, strategies as st
    from jpetra_b200.csrmatrix import CSRMatrix

    @settings(max_examples=int(__import__("os").environ.get("JP_HYPOTHESIS_EXAMPLES", "60")), deadline=None, derandomize=True)
    @given(st.integers(0, 2 ** 31), st.integers(1, 40), st.integers(0, 12), st.integers(1, 60))
    def run(seed, n_rows, max_len, n_global):
        rng = np.random.default_rng(seed)
        lens = rng.integers(0, max_len + 1, size=n_rows)
        ptr = np.zeros(n_rows + 1, dtype=np.int64)
        np.cumsum(lens, out=ptr[1:])
        gids = rng.integers(1, n_global + 1, size=int(ptr[-1])).astype(np.int64)
        vals = rng.standard_normal(len(gids))
        dom_n = int(rng.integers(0, n_global + 1))
        dom_min = int(rng.integers(1, n_global - dom_n + 2))
        got = emul.fill_complete(ptr, gids, vals, 1, n_global, dom_min, dom_n)
        # expected column map: used domain ids in order, then the other referenced ids ascending (= by owner, then id)
        used = np.unique(gids)
        local = used[(used >= dom_min) & (used < dom_min + dom_n)]
        remote = used[(used < dom_min) | (used >= dom_min + dom_n)]
        colmap = np.concatenate([local, remote])
        assert np.array_equal(got["colmap"], colmap) and got["n_local_used"] == len(local)
        n_same = 0
        while n_same < min(dom_n, len(colmap)) and colmap[n_same] == dom_min + n_same:
            n_same += 1
        assert got["n_same"] == n_same
        lid = {int(g): i for i, g in enumerate(colmap)}
        lcols = np.array([lid[int(g)] for g in gids], dtype=np.int64)
        eptr, ecols, evals = CSRMatrix._sortAndMerge(ptr, lcols, vals)
        assert np.array_equal(got["rowptr"], eptr) and np.array_equal(got["colind"], ecols) and np.array_equal(got["vals"], evals)

    run()
