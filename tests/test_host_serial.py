"""Host mirror (juliapetra.jl_b200) on SerialComm, CPU only: the reference's own serial golden vectors replayed
through the mirror's API, and bit-exact comparison of maps / column remapping / CSR with the CPU oracle.
No compute calls (no GPU here): fillComplete's host half is exercised through _fillCompleteHost."""
import numpy as np
import pytest

from helpers import O, jp, oracle_problem, problem_rows, global_rows


# ---- test/UtilsTests.jl:3-14 through the mirror --------------------------------------------------
def test_compute_offsets():
    rp = np.empty(30, dtype=np.int64)
    jp.computeOffsets(rp, 9)
    assert rp.tolist() == list(range(1, 9 * 30, 9))
    ne = np.arange(1, 30)
    assert jp.computeOffsets(rp, ne) == sum(range(1, 30))
    assert rp.tolist() == [int(ne[: i - 1].sum()) + 1 for i in range(1, 31)]
    for m in (30, 31, 42):
        with pytest.raises(jp.InvalidArgumentError):
            jp.computeOffsets(rp, np.zeros(m, dtype=np.int64))


# ---- test/BlockMapTests.jl:3-135 through the mirror ---------------------------------------------
def _serial_checks(m, m2, diff):
    for i in range(1, 6):
        assert m.myLID(i) and m.myGID(i) and m.lid(i) == i and m.gid(i) == i
    for bad in (-1, 0, 6, 30):
        assert not m.myLID(bad) and not m.myGID(bad) and m.lid(bad) == 0 and m.gid(bad) == 0
    assert not m.distributedGlobal()
    assert (m.numGlobalElements(), m.numMyElements()) == (5, 5)
    assert (m.minMyGID(), m.maxMyGID(), m.minAllGID(), m.maxAllGID(), m.minLID(), m.maxLID()) == (1, 5, 1, 5, 1, 5)
    p, l = m.remoteIDList([1, 2, 3, 4, 5])
    assert p.tolist() == [1] * 5 and l.tolist() == [1, 2, 3, 4, 5]
    assert m.myGlobalElements().tolist() == [1, 2, 3, 4, 5]
    assert m.sameAs(m2) and m2.sameAs(m) and not m.sameAs(diff) and not diff.sameAs(m)
    assert m.linearMap() and m.uniqueGIDs()
    assert m.getComm() == jp.SerialComm()


def test_blockmap_serial():
    c = jp.SerialComm()
    for bad in (-8, -1):
        with pytest.raises(jp.InvalidArgumentError):
            jp.BlockMap(bad, c)
    jp.BlockMap(0, c), jp.BlockMap(1, c)
    _serial_checks(jp.BlockMap(5, c), jp.BlockMap(5, c), jp.BlockMap(6, c))
    for N, n in ((-8, 4), (-2, 4), (5, -6), (4, -1)):
        with pytest.raises(jp.InvalidArgumentError):
            jp.BlockMap(N, n, c)
    jp.BlockMap(0, 0, c), jp.BlockMap(1, 1, c)
    _serial_checks(jp.BlockMap(5, 5, c), jp.BlockMap(5, 5, c), jp.BlockMap(6, 6, c))
    _serial_checks(jp.BlockMap(-1, 5, c), jp.BlockMap(-1, 5, c), jp.BlockMap(-1, 6, c))
    jp.BlockMap(0, [], c), jp.BlockMap(1, [1], c)
    _serial_checks(jp.BlockMap(5, [1, 2, 3, 4, 5], c), jp.BlockMap(5, [1, 2, 3, 4, 5], c), jp.BlockMap(6, [1, 2, 3, 4, 5, 6], c))


def test_blockmap_nonlinear_lookup_matches_oracle():
    rng = np.random.default_rng(1)
    gids = np.concatenate([np.arange(10, 20), rng.permutation(np.arange(100, 160))[:40], [12, 300]])  # prefix + scattered + a duplicate
    m = jp.BlockMap(-1, gids, jp.SerialComm())
    om = O.OMap.from_gids(O.OComm(1), -1, [gids])
    q = np.arange(0, 320)
    assert m.lid(q).tolist() == [om.lid(1, int(g)) for g in q]
    assert m.gid(np.arange(0, 60)).tolist() == [om.gid(1, int(l)) for l in range(0, 60)]


# ---- test/SerialCommTests.jl through the mirror ---------------------------------------------------
def test_serial_comm_and_distributor():
    c = jp.SerialComm()
    assert c.myPid() == 1 and c.numProc() == 1
    assert c.sumAll([3, 4]).tolist() == [3, 4] and c.gatherAll([1.5]).tolist() == [1.5]
    with pytest.raises(jp.InvalidArgumentError):
        c.broadcastAll([1], 2)
    d = c.createDistributor()
    for bad in ([1, 1, 1, 2], [1, 1, 2, 1], [2, 1, 1]):
        with pytest.raises(jp.InvalidArgumentError):
            d.createFromSends(bad)
    assert [d.createFromSends([1] * n) for n in (1, 2, 5, 8)] == [1, 2, 5, 8]
    with pytest.raises(jp.InvalidArgumentError):
        d.createFromRecvs([1, 2, 3, 4], [1, 1, 1, 2])
    g, p = d.createFromRecvs([2, 3, 4, 5], [1, 1, 1, 1])
    assert g.tolist() == [2, 3, 4, 5] and p.tolist() == [1, 1, 1, 1]
    assert d.resolve([3, 4, 5]) == [3, 4, 5] and d.resolveReverse(["a", "k"]) == ["a", "k"]
    with pytest.raises(jp.InvalidStateError):
        d.resolveWaits()
    with pytest.raises(jp.InvalidStateError):
        d.resolveReverseWaits()
    d.resolvePosts([1, 2, 3])
    d.resolvePosts([6, 7, 8, 9])
    assert d.resolveWaits() == [6, 7, 8, 9]
    with pytest.raises(jp.InvalidStateError):
        d.resolveWaits()


# ---- test/Import-Export Tests.jl:9-26 -------------------------------------------------------------
@pytest.mark.parametrize("cls", [jp.Import, jp.Export])
def test_trivial_plan(cls):
    c = jp.SerialComm()
    src, des = jp.BlockMap(8, 8, c), jp.BlockMap(8, 8, c)
    plan = cls(src, des)
    assert plan.numSameIDs() == 8 and plan.isLocallyComplete()
    for f in ("permuteToLIDs", "permuteFromLIDs", "remoteLIDs", "exportLIDs", "exportPIDs"):
        assert len(getattr(plan, f)()) == 0
    assert plan.sourceMap() is src and plan.targetMap() is des


# ---- test/CSRMatrixTests.jl:120-138 (fillComplete result), host half -------------------------------------
def test_csrmatrix_fill_golden():
    c = jp.SerialComm()
    m = jp.BlockMap(2, 2, c)
    A = jp.CSRMatrix(m, 2, jp.STATIC_PROFILE)
    A.insertGlobalValues(1, [1, 2], [2, 3])
    A.insertGlobalValues(2, [1, 2], [5, 7])
    with pytest.raises(jp.InvalidArgumentError):  # src/CSRMatrix.jl:644-648
        A.insertGlobalValues(1, [1], [1.0])
    with pytest.raises(jp.InvalidStateError):  # src/RowMatrix.jl:267-269
        A.apply_(None, None)
    A._fillCompleteHost()
    assert A.localIndices1D.tolist() == [1, 2, 1, 2]
    assert A.rowOffsets.tolist() == [1, 3, 5]
    assert (A.getLocalRowCopy(1)[0].tolist(), A.getLocalRowCopy(1)[1].tolist()) == ([1, 2], [2, 3])
    assert (A.getLocalRowCopy(2)[0].tolist(), A.getLocalRowCopy(2)[1].tolist()) == ([1, 2], [5, 7])
    assert A.colMap is m and A.importer is None and A.exporter is None
    with pytest.raises(jp.InvalidStateError):  # src/CSRMatrix.jl:709-712
        A._fillCompleteHost()


# ---- bit-exact CSR / column map against the oracle, serial -------------------------------------------------
@pytest.mark.parametrize("kind,N", [("5pt", 13), ("7pt", 7), ("27pt", 5), ("powerlaw", 700)])
def test_fill_complete_matches_oracle_serial(kind, N):
    from jpetra_b200 import synthetic
    c = jp.SerialComm()
    m = jp.BlockMap(global_rows(kind, N), c)
    A = synthetic.build_matrix(jp, m, kind, N, chunk_rows=97, fill=False, max_len=200)
    A._fillCompleteHost(m, m)
    _, _, OA = oracle_problem(kind, N, 1, max_len=200)
    rp, ci, va = OA.csr(1)
    assert np.array_equal(A.rowOffsets, rp)
    assert np.array_equal(A.localIndices1D, ci)
    assert np.array_equal(A.values, va)  # bit-exact, including duplicate merging order
    assert np.array_equal(A.colMap.myGlobalElements(), OA.col_map.my_gids(1))
    assert (A.importer is None) == (OA.importer is None)


def test_sort_and_merge_unsorted_duplicates():
    # unsorted insertion with duplicates in several rows; one column unused -> column map != domain map in serial
    c = jp.SerialComm()
    m = jp.BlockMap(6, c)
    rows = {1: ([3, 1, 3, 2, 1], [1.0, 2.0, 0.25, 4.0, 1e-3]), 2: ([2], [5.0]), 3: ([6, 5, 6, 6], [1.0, 2.0, 3.0, 4.5]), 4: ([], []),
            5: ([1, 2, 3], [1.0, 1.0, 1.0]), 6: ([6, 1], [7.0, 8.0])}
    A = jp.CSRMatrix(m, [5, 1, 4, 0, 3, 2], jp.STATIC_PROFILE)
    oc = O.OComm(1)
    om = O.OMap.uniform(oc, 6)
    OA = O.OMatrix(om, [[5, 1, 4, 0, 3, 2]])
    for g, (ci, va) in rows.items():
        if ci:
            A.insertGlobalValues(g, ci, va)
            OA.insert(1, g, ci, va)
    A._fillCompleteHost(m, m)
    OA.fill_complete(om, om)
    rp, ci, va = OA.csr(1)
    assert np.array_equal(A.rowOffsets, rp) and np.array_equal(A.localIndices1D, ci) and np.array_equal(A.values, va)
    assert A.colMap.myGlobalElements().tolist() == OA.col_map.my_gids(1).tolist() == [1, 2, 3, 5, 6]
    imp, oimp = A.importer, OA.importer
    assert imp is not None and oimp is not None
    assert imp.numSameIDs() == oimp.field(1, "numSameIDs") == 3
    assert imp.permuteToLIDs().tolist() == oimp.field(1, "permuteToLIDs").tolist() == [4, 5]
    assert imp.permuteFromLIDs().tolist() == oimp.field(1, "permuteFromLIDs").tolist() == [5, 6]


def test_c_abi_exports_every_declared_symbol():
    """The C-ABI library loads on a CPU-only box and exports every function include/jpetra_b200.h declares."""
    import os
    import re
    from jpetra_b200 import _lib
    lib = _lib.load()
    header = open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "include", "jpetra_b200.h")).read()
    declared = set(re.findall(r"\bint32_t\s+(jp_\w+)\s*\(", header))
    assert len(declared) >= 50
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in the header but not exported"
    assert declared == set(_lib.PROTOTYPES), declared ^ set(_lib.PROTOTYPES)
    assert lib.jp_version() == 1


def test_no_cpu_fallback():
    """Without a device every compute entry fails loudly (there is no CPU path to fall back to)."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(jp.BackendError):
        jp.DenseMultiVector(jp.BlockMap(4, jp.SerialComm()), 1)
    m = jp.BlockMap(4, jp.SerialComm())
    A = jp.CSRMatrix(m, 1, jp.STATIC_PROFILE)
    for g in range(1, 5):
        A.insertGlobalValues(g, [g], [1.0])
    with pytest.raises(jp.BackendError):
        A.fillComplete(device=True)  # the device fillComplete does not quietly take the host path either
    assert A.isFillActive()


# ---- committed golden fixtures (tests/golden/fixtures.json, generated from the oracle by make_fixtures.py) ----------
def _fixtures():
    import json
    import os
    return json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "fixtures.json")))


def _digest(a):
    import hashlib
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def test_oracle_reproduces_committed_fixtures():
    fx = _fixtures()
    P, N = 4, 6
    _, m, A = oracle_problem("7pt", N, P)
    for pid in range(1, P + 1):
        g = fx["7pt_6_P4"][str(pid)]
        rp, ci, va = A.csr(pid)
        assert (_digest(rp), _digest(ci), _digest(va)) == (g["rowptr_sha256"], g["colind_sha256"], g["vals_sha256"])
        assert _digest(A.col_map.my_gids(pid)) == g["colMap_sha256"]
        assert _digest(A.importer.field(pid, "remoteLIDs")) == g["remoteLIDs_sha256"]
        assert _digest(A.importer.field(pid, "exportLIDs")) == g["exportLIDs_sha256"]
        assert A.importer.distributor.field(pid, "procs_to").tolist() == g["procs_to"]


@pytest.mark.parametrize("key,kind,size", [("7pt_64_serial", "7pt", 64), ("5pt_300_serial", "5pt", 300), ("27pt_20_serial", "27pt", 20)])
def test_host_mirror_csr_feeds_the_committed_apply_digest(key, kind, size):
    """the host mirror's CSR, pushed through the oracle's localApply, reproduces the committed digests (CPU check of
    the inputs the GPU test consumes)"""
    from jpetra_b200 import synthetic
    g = _fixtures()[key]
    c = jp.SerialComm()
    m = jp.BlockMap(global_rows(kind, size), c)
    A = synthetic.build_matrix(jp, m, kind, size, fill=False)
    A._fillCompleteHost(m, m)
    n = m.numMyElements()
    assert (n, len(A.values)) == (g["n"], g["nnz"])
    x = synthetic.vector_block(1, n, 1, seed=1)
    y0 = synthetic.vector_block(1, n, 1, seed=2)
    assert (_digest(x), _digest(y0)) == (g["x_sha256"], g["y0_sha256"])
    y = O.local_apply_raw(y0.copy(order="F"), A.rowOffsets, A.localIndices1D, A.values, x, O.NO_TRANS, 3.0, 0.5)
    assert _digest(y) == g["apply_3_05_sha256"]
