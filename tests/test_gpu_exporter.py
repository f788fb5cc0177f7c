"""apply! when rangeMap != rowMap (jp_apply_export) on the GPU against the CPU oracle's exporter
branch (itself checked against a dense operator in tests/test_oracle_golden.py).

The device path is host orchestration of kernels the importer / transposed paths already use; the multi-rank case lives in
tests/multi_gpu_cases.py."""
import numpy as np
import pytest

from helpers import O, jp

pytestmark = pytest.mark.gpu


def _serial_permuted_problem(n, seed):
    """one rank; the row map lists the n rows in a shuffled order, domain = range = the uniform map: the Export
    (rowMap -> rangeMap) has no remote entries but a non-trivial same / permute split"""
    rng = np.random.default_rng(seed)
    order = np.arange(1, n + 1)
    order[n // 3:] = rng.permutation(order[n // 3:])  # a same-ID prefix, then permuted ids
    lens = rng.integers(1, 7, size=n)
    rp = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(lens, out=rp[1:])
    cols = np.concatenate([np.concatenate([[g], rng.integers(1, n + 1, size=l - 1)]) for g, l in zip(order, lens)]).astype(np.int64)
    vals = rng.standard_normal(len(cols))
    return order, rp, cols, vals


@pytest.mark.parametrize("n,k", [(40, 1), (1000, 3)])
def test_exporter_branch_serial_permuted_row_map(n, k):
    order, rp, cols, vals = _serial_permuted_problem(n, seed=n)
    comm = jp.SerialComm()
    uni = jp.BlockMap(n, comm)
    rowmap = jp.BlockMap(-1, order, comm)
    A = jp.CSRMatrix(rowmap, np.diff(rp), jp.STATIC_PROFILE)
    A.insertGlobalValuesBatch(order, rp, cols, vals)
    A.fillComplete(uni, uni)
    assert A.getExporter() is not None
    oc = O.OComm(1)
    ouni, orow = O.OMap.uniform(oc, n), O.OMap.from_gids(oc, -1, [order])
    OA = O.OMatrix(orow, [np.diff(rp)])
    OA.insert_rows(1, order, rp, cols, vals)
    OA.fill_complete(ouni, ouni)
    rng = np.random.default_rng(3)
    x, y0 = rng.standard_normal((n, k)), rng.standard_normal((n, k))
    for mode, omode in ((jp.NO_TRANS, O.NO_TRANS), (jp.TRANS, O.TRANS)):
        for alpha, beta in ((3.0, 0.5), (1.0, 0.0), (0.0, 2.0)):
            OX, OY = O.OMV.from_arrays(ouni, [x]), O.OMV.from_arrays(ouni, [y0])
            OA.apply(OY, OX, omode, alpha, beta)
            X, Y = jp.DenseMultiVector(uni, x), jp.DenseMultiVector(uni, y0)
            A.apply_(Y, X, mode, alpha, beta)
            ref = OY.get(1)
            assert np.abs(Y.data - ref).max() <= 1e-12 * 50 * (np.abs(ref).max() + 1.0), (mode, alpha, beta)
    with pytest.raises(jp.InvalidStateError):
        A.apply_host(x)
