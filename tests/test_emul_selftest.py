"""Self-tests of the CPU execution model (tests/cuda_emul): a kernel with a missing barrier must be exposed by the
alternative thread schedules, a correct one must not care; and the product's kernels give bit-identical results under
forward, reverse and random schedules (no result depends on the order in which threads reach a barrier)."""
import numpy as np
import pytest

import emul
from helpers import O, oracle_problem, synthetic


@pytest.fixture(autouse=True)
def _restore_schedule():
    yield
    emul.set_schedule("forward")


def test_missing_barrier_is_exposed_and_correct_kernel_is_not():
    n = 64
    expect = np.array([1000 + (t + 1) % n for t in range(n)])
    wrong = []
    for mode, seed in (("forward", 0), ("reverse", 0), ("random", 3)):
        emul.set_schedule(mode, seed)
        assert np.array_equal(emul.selftest_neighbour(n, True), expect)
        wrong.append(int((emul.selftest_neighbour(n, False) != expect).sum()))
    assert wrong[0] == n - 1 and wrong[1] == 1 and 0 < wrong[2] < n  # each schedule shows the race differently


def test_kernels_do_not_depend_on_the_thread_schedule():
    c, m, A = oracle_problem("27pt", 6, 1)
    rp, ci, va = A.csr(1)
    rp, ci = rp - 1, ci - 1
    n = len(rp) - 1
    x1, y1 = synthetic.vector_block(1, n, 1, seed=1), synthetic.vector_block(1, n, 1, seed=2)
    x9, y9 = synthetic.vector_block(1, n, 9, seed=1), synthetic.vector_block(1, n, 9, seed=2)
    c2, m2, A2 = oracle_problem("powerlaw", 1500, 2, max_len=1200)
    info = m2.info(2)
    from helpers import problem_rows
    rows2 = problem_rows("powerlaw", 1500, info["minMyGID"], info["maxMyGID"], max_len=1200)

    def everything():
        out = [emul.local_apply(rp, ci, va, x1, y1, 3.0, 0.5)[0]]                                  # spmv_kernel
        out += [emul.local_apply(rp, ci, va, x9, y9, 3.0, 0.5, variant=v, cap=cap)[0] for v, cap in ((0, 2048), (6, 3456))]
        out += [emul.spmm_dia(rp, ci, va, x9, y9, 3.0, 0.5)[0]]                                     # spmm_dia_kernel
        yd, dots = emul.spmv_dot(rp, ci, va, x1, x1, y1, 3.0, 0.5)
        out += [yd, dots]
        ya, sums = emul.axpby_norm2(y9, -0.37, x9, 1.0, blocks=3, repeats=2)
        out += [ya, sums, emul.mv_reduce(0, x9, y9, blocks=3)]
        f = emul.fill_complete(rows2[1], rows2[2], rows2[3], 1, 1500, info["minMyGID"], info["numMyElements"])
        out += [f["rowptr"], f["colind"], f["vals"], f["colmap"]]
        return out

    emul.set_schedule("forward")
    base = everything()
    for mode, seed in (("reverse", 0), ("random", 11)):
        emul.set_schedule(mode, seed)
        for a, b in zip(base, everything()):
            assert np.array_equal(a, b), mode
