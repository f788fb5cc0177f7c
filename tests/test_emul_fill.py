"""Device fillComplete (csrc/jp_fill_kernels.cuh + jp_fill_pipeline.h) executed on the CPU model of tests/cuda_emul and
compared BIT-EXACTLY with the CPU oracle's fillComplete (column map, local indices, merged values, row pointers,
numSameIDs) for 1-4 simulated ranks.  The same cases run on the GPU in tests/test_gpu_fill.py."""
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
