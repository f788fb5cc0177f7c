"""SPMD bodies of the multi-process CPU tests (gloo).  Each function runs on every rank of a torch.distributed
group and compares THIS rank's host-mirror results with the CPU oracle's simulation of all ranks."""
import os
import sys
import traceback

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def _entry(rank, world, port, case, args):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        globals()[case](rank, world, *args)
    except BaseException:
        traceback.print_exc()
        raise
    finally:
        dist.destroy_process_group()


def run(case, world, *args):
    import torch.multiprocessing as mp
    from helpers import free_port
    mp.start_processes(_entry, args=(world, free_port(), case, args), nprocs=world, join=True, start_method="spawn")


# ---- test/MPICommTests.jl:13-65 ----------------------------------------------------------------------------
def case_comm_and_distributor(rank, world):
    import jpetra_b200 as jp
    comm = jp.TorchDistComm()
    pid, P = comm.myPid(), comm.numProc()
    assert P == world and pid == rank + 1
    for root, v in ((1, 8), (2, 5)):
        assert comm.broadcastAll([pid, v], root).tolist() == [root, v]
    assert comm.gatherAll([pid]).tolist() == list(range(1, P + 1))
    assert comm.gatherAll([pid, pid * 2, pid * 3]).tolist() == [x for p in range(1, P + 1) for x in (p, 2 * p, 3 * p)]
    comm.barrier()
    tot = P * (P + 1) // 2
    assert comm.sumAll([pid]).tolist() == [tot]
    assert comm.sumAll([8, 3, pid, 2]).tolist() == [8 * P, 3 * P, tot, 2 * P]
    assert comm.maxAll([pid, -pid, 8]).tolist() == [P, -1, 8]
    assert comm.minAll([pid, -pid, 6]).tolist() == [1, -P, 6]
    assert comm.scanSum([pid]).tolist() == [pid * (pid + 1) // 2]
    assert comm.scanSum([5, -2 * pid, 3]).tolist() == [pid * 5, sum(range(-2, -2 * pid - 1, -2)), pid * 3]
    dist = comm.createDistributor()
    for f in (dist.resolveWaits, dist.resolveReverseWaits):
        try:
            f()
            raise AssertionError("expected InvalidStateError")
        except jp.InvalidStateError:
            pass
    everyone = list(range(1, P + 1))
    assert dist.createFromSends(everyone) == P
    dist.resolvePosts([pid * j for j in everyone])
    assert dist.resolveWaits().tolist() == [pid * j for j in everyone]
    try:
        dist.resolveWaits()
        raise AssertionError("expected InvalidStateError")
    except jp.InvalidStateError:
        pass
    # not blocked by processor (:57-62)
    assert dist.createFromSends(everyone + everyone) == 2 * P
    got = dist.resolve([P * 5 * 0 + 5 * t + pid for t in range(2 * P)])
    exp = []
    for j in everyone:
        exp += [(pid - 1) * 5 + j, (pid - 1 + P) * 5 + j]
    assert got.tolist() == exp, (got, exp)
    # createFromRecvs (:65)
    eg, ep = dist.createFromRecvs([pid + 5 * t for t in range(P)], everyone)
    assert eg.tolist() == [(pid - 1) * 5 + j for j in everyone] and ep.tolist() == everyone
    # plan fields against the oracle
    import oracle as O
    od = O.ODist(O.OComm(P))
    od.create_from_recvs([[p + 5 * t for t in range(P)] for p in everyone], [everyone] * P)
    f = dist.plan_fields()
    for name in ("procs_to", "lengths_to", "starts_to", "indices_to", "procs_from", "lengths_from", "starts_from"):
        assert np.array_equal(f[name], od.field(pid, name)), name
    s = od.scalars(pid)
    assert (dist.numSends, dist.numRecvs, dist.selfMsg, dist.totalRecvLength) == (s["numSends"], s["numRecvs"], s["selfMsg"], s["totalRecvLength"])


# ---- test/MPIBlockMapTests.jl ----------------------------------------------------------------------------------
def case_blockmaps(rank, world):
    import jpetra_b200 as jp
    import oracle as O
    comm = jp.TorchDistComm()
    pid, P = comm.myPid(), comm.numProc()
    oc = O.OComm(P)
    N = 5 * P
    maps = [(jp.BlockMap(N, comm), O.OMap.uniform(oc, N)),
            (jp.BlockMap(N, 5, comm), O.OMap.linear(oc, N, [5] * P)),
            (jp.BlockMap(N, np.arange(1, 6) + 5 * (pid - 1), comm), O.OMap.from_gids(oc, N, [np.arange(1, 6) + 5 * (p - 1) for p in range(1, P + 1)])),
            (jp.BlockMap(N + 3, comm), O.OMap.uniform(oc, N + 3)),  # uneven split
            (jp.BlockMap(-1, 3 + pid, comm), O.OMap.linear(oc, -1, [3 + p for p in range(1, P + 1)]))]
    for m, om in maps:
        info = om.info(pid)
        assert (m.numMyElements(), m.minMyGID(), m.maxMyGID(), m.minAllGID(), m.maxAllGID(), m.minLID(), m.maxLID(), int(m.linearMap()),
                int(m.distributedGlobal()), m.numGlobalElements()) == tuple(info[k] for k in O.OMap.INFO)
        q = np.arange(-1, m.numGlobalElements() + 3)
        assert m.lid(q).tolist() == [om.lid(pid, int(g)) for g in q]
        assert m.gid(np.arange(-1, 12)).tolist() == [om.gid(pid, int(l)) for l in range(-1, 12)]
        allg = np.arange(1, m.numGlobalElements() + 1)
        pr, lr = m.remoteIDList(allg)
        opr, olr = om.remote_id_list(pid, allg)
        assert np.array_equal(pr, opr) and np.array_equal(lr, olr)
        assert np.array_equal(m.myGlobalElements(), om.my_gids(pid))
    m = maps[0][0]
    assert m.lid(np.arange(1, N + 1)).tolist() == [(i - 1) % 5 + 1 if -(-i // 5) == pid else 0 for i in range(1, N + 1)]
    pr, lr = m.remoteIDList(np.arange(1, N + 1))  # test/MPIBlockMapTests.jl:60-61
    assert pr.tolist() == np.repeat(np.arange(1, P + 1), 5).tolist() and lr.tolist() == list(range(1, 6)) * P
    assert maps[0][0].sameAs(maps[1][0]) and maps[0][0].sameAs(maps[2][0]) and not maps[0][0].sameAs(maps[3][0])


# ---- Import / Export lists and fillComplete layout, bit-exact against the oracle ----------------------------------
def case_shifted_plans(rank, world):
    import jpetra_b200 as jp
    import oracle as O
    from helpers import assert_plan_equal
    comm = jp.TorchDistComm()
    pid, P = comm.myPid(), comm.numProc()
    n = 4
    src = jp.BlockMap(P * n, comm)
    des = jp.BlockMap(n * P, np.arange(1, n + 1) + n * (pid % P), comm)  # test/MPIimport-export Tests.jl:3-4
    oc = O.OComm(P)
    osrc = O.OMap.uniform(oc, P * n)
    odes = O.OMap.from_gids(oc, n * P, [np.arange(1, n + 1) + n * (p % P) for p in range(1, P + 1)])
    for cls, make in ((jp.Import, O.OPlan.make_import), (jp.Export, O.OPlan.make_export)):
        plan = cls(src, des)
        assert plan.numSameIDs() == 0 and len(plan.permuteToLIDs()) == 0 and plan.isLocallyComplete()
        assert_plan_equal(plan, make(osrc, odes), pid)


def case_fill_complete(rank, world, kind, N, max_len):
    import jpetra_b200 as jp
    from jpetra_b200 import synthetic
    from helpers import assert_plan_equal, global_rows, oracle_problem
    comm = jp.TorchDistComm()
    pid, P = comm.myPid(), comm.numProc()
    m = jp.BlockMap(global_rows(kind, N), comm)
    A = synthetic.build_matrix(jp, m, kind, N, chunk_rows=1000, fill=False, max_len=max_len)
    A._fillCompleteHost(m, m)
    _, _, OA = oracle_problem(kind, N, P, max_len=max_len)
    rp, ci, va = OA.csr(pid)
    assert np.array_equal(A.rowOffsets, rp)
    assert np.array_equal(A.localIndices1D, ci)
    assert np.array_equal(A.values, va)
    assert np.array_equal(A.colMap.myGlobalElements(), OA.col_map.my_gids(pid))
    assert (A.importer is None) == (OA.importer is None)
    if A.importer is not None:
        assert_plan_equal(A.importer, OA.importer, pid)
    assert A.exporter is None and OA.exporter is None


def case_laplacian_1d_golden(rank, world):
    """test/CSRMatrixMPITests.jl layout through the mirror, world=4"""
    import jpetra_b200 as jp
    comm = jp.TorchDistComm()
    pid, P = comm.myPid(), comm.numProc()
    n = 2
    N = P * n
    m = jp.BlockMap(N, comm)
    g = m.myGlobalElements()
    A = jp.CSRMatrix(m, [2 if x in (1, N) else 3 for x in g], jp.STATIC_PROFILE)
    for x in g:
        x = int(x)
        idx = [2] if x == 1 else [N - 2] if x == N else [x - 1, x + 1]
        A.insertGlobalValues(x, idx, [-1.0, -1.0][: len(idx)])
        A.insertGlobalValues(x, [x], [2.0])
    A._fillCompleteHost(m, m)
    exp = {1: ([1, 2, 3], [3], [2], [2], [1, 3, 6], [1, 2, 1, 2, 3]), 2: ([3, 4, 2, 5], [3, 4], [1, 2], [1, 3], [1, 4, 7], [1, 2, 3, 1, 2, 4]),
           3: ([5, 6, 4, 7], [3, 4], [1, 2], [2, 4], [1, 4, 7], [1, 2, 3, 1, 2, 4]), 4: ([7, 8, 6], [3], [1], [3], [1, 4, 6], [1, 2, 3, 2, 3])}[pid]
    assert A.colMap.myGlobalElements().tolist() == exp[0]
    assert A.importer.remoteLIDs().tolist() == exp[1] and A.importer.exportLIDs().tolist() == exp[2]
    assert A.importer.exportPIDs().tolist() == exp[3] and A.importer.numSameIDs() == 2
    assert A.rowOffsets.tolist() == exp[4] and A.localIndices1D.tolist() == exp[5]


def case_irregular_matrix(rank, world):
    """more rows than one rank can fill (an empty rank), unused local columns (permutes) and remotes at once:
    maps, column map, Import lists and CSR bit-exact against the oracle"""
    import jpetra_b200 as jp
    import oracle as O
    from helpers import assert_plan_equal
    comm = jp.TorchDistComm()
    pid, P = comm.myPid(), comm.numProc()
    N = P + 1 if P > 2 else 5  # uneven split; with P=4: rows per rank 2,1,1,1
    rng = np.random.default_rng(11)
    cols_of = {g: sorted(set(rng.integers(1, N + 1, size=3).tolist()) - {2}) or [1] for g in range(1, N + 1)}  # GID 2 is never a column
    vals_of = {g: rng.standard_normal(len(cols_of[g])) for g in range(1, N + 1)}
    m = jp.BlockMap(N, comm)
    oc = O.OComm(P)
    om = O.OMap.uniform(oc, N)
    mine = m.myGlobalElements()
    A = jp.CSRMatrix(m, [len(cols_of[int(g)]) for g in mine], jp.STATIC_PROFILE)
    for g in mine:
        A.insertGlobalValues(int(g), cols_of[int(g)], vals_of[int(g)])
    A._fillCompleteHost(m, m)
    OA = O.OMatrix(om, [[len(cols_of[int(g)]) for g in om.my_gids(p)] for p in range(1, P + 1)])
    for p in range(1, P + 1):
        for g in om.my_gids(p):
            OA.insert(p, int(g), cols_of[int(g)], vals_of[int(g)])
    OA.fill_complete(om, om)
    rp, ci, va = OA.csr(pid)
    assert np.array_equal(A.rowOffsets, rp) and np.array_equal(A.localIndices1D, ci) and np.array_equal(A.values, va)
    assert np.array_equal(A.colMap.myGlobalElements(), OA.col_map.my_gids(pid))
    assert (A.importer is None) == (OA.importer is None)
    if A.importer is not None:
        assert_plan_equal(A.importer, OA.importer, pid)


def case_export_plan(rank, world):
    """a non-trivial Export (source = overlapping GID lists, target = uniform map): lists bit-exact against the oracle"""
    import jpetra_b200 as jp
    import oracle as O
    from helpers import assert_plan_equal
    comm = jp.TorchDistComm()
    pid, P = comm.myPid(), comm.numProc()
    n = 3
    N = n * P
    tgt = jp.BlockMap(N, comm)
    # every rank holds its own rows plus the first row of the next rank (an overlapping, non-linear source map)
    def src_gids(p):
        own = list(range((p - 1) * n + 1, p * n + 1))
        return own + [(p % P) * n + 1]
    src = jp.BlockMap(-1, src_gids(pid), comm)
    oc = O.OComm(P)
    otgt = O.OMap.uniform(oc, N)
    osrc = O.OMap.from_gids(oc, -1, [src_gids(p) for p in range(1, P + 1)])
    assert src.numGlobalElements() == osrc.info(pid)["numGlobalElements"] == N
    assert_plan_equal(jp.Export(src, tgt), O.OPlan.make_export(osrc, otgt), pid)


def case_device_fill_glue(rank, world, kind, N, max_len):
    """CSRMatrix.fillComplete(device=True) on every rank, with the four C-ABI calls behind it served by the CPU execution
    model of the fillComplete kernels (tests/emul.py): the multi-rank glue (error consensus, column-map BlockMap, Import)
    and the kernels' results against the oracle and against the host path"""
    import emul
    import jpetra_b200 as jp
    from jpetra_b200 import csrmatrix, synthetic
    from helpers import assert_plan_equal, global_rows, oracle_problem
    store = {}

    def fake_fill(n_rows, ptr0, gcols, vals, gmin, gmax, dom_min, n_dom):
        try:
            out = emul.fill_complete(ptr0, gcols, vals, gmin, gmax, dom_min, n_dom)
        except emul.EmulError as e:
            raise jp.InvalidArgumentError(str(e))
        h = len(store) + 1
        store[h] = out
        return h, [out["n_cols"], out["n_local_used"], out["n_same"], out["nnz"]]

    csrmatrix._device_fill_complete = fake_fill
    csrmatrix._device_col_map = lambda h, n_cols: store[h]["colmap"]
    csrmatrix._device_csr_download = lambda h, n_rows, nnz: (store[h]["rowptr"], store[h]["colind"], store[h]["vals"])
    csrmatrix._device_csr_destroy = lambda h: store.pop(h, None)
    jp.Import.device_plan = lambda self: 0
    comm = jp.TorchDistComm()
    pid, P = comm.myPid(), comm.numProc()
    m = jp.BlockMap(global_rows(kind, N), comm)
    A = synthetic.build_matrix(jp, m, kind, N, chunk_rows=1000, fill=False, max_len=max_len)
    A.fillComplete(m, m, device=True)
    assert A.isFillComplete() and A.h in store  # the device path was taken
    Ah = synthetic.build_matrix(jp, m, kind, N, chunk_rows=1000, fill=False, max_len=max_len)
    Ah._fillCompleteHost(m, m)
    _, _, OA = oracle_problem(kind, N, P, max_len=max_len)
    rp, ci, va = OA.csr(pid)
    for B in (A, Ah):
        assert np.array_equal(B.rowOffsets, rp) and np.array_equal(B.localIndices1D, ci) and np.array_equal(B.values, va)
        assert np.array_equal(B.colMap.myGlobalElements(), OA.col_map.my_gids(pid))
        assert (B.importer is None) == (OA.importer is None)
        if B.importer is not None:
            assert_plan_equal(B.importer, OA.importer, pid)
    assert A.getLocalNumEntries() == Ah.getLocalNumEntries() == len(va)
    # a column GID nobody owns: every rank raises, whichever rank saw it (maxAll of the error flag)
    B = jp.CSRMatrix(m, 2, jp.STATIC_PROFILE)
    for g in m.myGlobalElements():
        B.insertGlobalValues(int(g), [int(g), int(g) if pid != P else m.numGlobalElements() + 5], [1.0, 2.0])
    try:
        B.fillComplete(m, m, device=True)
        raise AssertionError("expected makeColMap to report an error")
    except RuntimeError as e:
        assert "makeColMap" in str(e)


def case_overlapping_rows_fill(rank, world):
    """a matrix assembled on an overlapping row map with domain = range = the uniform map (rangeMap != rowMap, SURVEY.md
    8f item 4): column map, Import AND Export lists, CSR bit-exact against the oracle"""
    import jpetra_b200 as jp
    from fill_cases import overlapping_rows_problem
    from helpers import assert_plan_equal
    comm = jp.TorchDistComm()
    pid, P = comm.myPid(), comm.numProc()
    c, ouni, orow, OA, per_rank, dense = overlapping_rows_problem(P, 6)
    rows, rp, cols, vals = per_rank[pid - 1]
    uni = jp.BlockMap(P * 6, comm)
    rowmap = jp.BlockMap(-1, rows, comm)
    assert rowmap.numGlobalElements() == orow.info(pid)["numGlobalElements"]
    A = jp.CSRMatrix(rowmap, np.diff(rp), jp.STATIC_PROFILE)
    A.insertGlobalValuesBatch(rows, rp, cols, vals)
    A._fillCompleteHost(uni, uni)
    orp, oci, ova = OA.csr(pid)
    assert np.array_equal(A.rowOffsets, orp) and np.array_equal(A.localIndices1D, oci) and np.array_equal(A.values, ova)
    assert np.array_equal(A.colMap.myGlobalElements(), OA.col_map.my_gids(pid))
    assert A.exporter is not None and OA.exporter is not None
    assert_plan_equal(A.exporter, OA.exporter, pid)
    assert (A.importer is None) == (OA.importer is None)
    if A.importer is not None:
        assert_plan_equal(A.importer, OA.importer, pid)
