"""SPMD body of the multi-GPU parity test: launched by torchrun, one rank per GPU.
   torchrun --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tests/multi_gpu_cases.py
Every rank compares ITS part of the result with the CPU oracle's simulation of all N ranks."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

TOL = 1e-12


def dense_abs_bound(triples, N, xs, ys, trans, lo, hi):
    """rows [lo, hi) (0-based global ids) of 3*|A|*|x| + 0.5*|y0| (or |A|^T) for a small matrix given as (row GID, col GID,
    value) triples: the component-wise bound of the single-GPU tests, formed densely"""
    absA = np.zeros((N, N))
    for r, c, v in triples:
        absA[r - 1, c - 1] += abs(v)
    x, y0 = np.abs(np.vstack(xs)), np.abs(np.vstack(ys))
    b = 3.0 * ((absA.T if trans else absA) @ x) + 0.5 * y0
    return b[lo:hi]


def main():
    os.environ["JP_HOST_PIPELINE_MIN_BYTES"] = "1"  # apply_host takes the chunked three-stream pipeline even at test sizes
    import torch.distributed as dist
    dist.init_process_group("gloo")
    import jpetra_b200 as jp
    from jpetra_b200 import _lib, synthetic
    import oracle as O
    from helpers import assert_plan_equal, global_rows, oracle_problem

    rank, world = dist.get_rank(), dist.get_world_size()
    _lib.init(int(os.environ.get("LOCAL_RANK", rank)))
    comm = jp.NCCLComm.fromTorchDistributed()
    pid, P = comm.myPid(), comm.numProc()
    assert (pid, P) == (rank + 1, world)

    # ---- Comm collectives through the C ABI: test/MPICommTests.jl:13-36 ----
    everyone = list(range(1, P + 1))
    tot = P * (P + 1) // 2
    assert comm.broadcastAll([pid, 8], 1).tolist() == [1, 8]
    assert comm.broadcastAll([pid, 5], P).tolist() == [P, 5]
    assert comm.gatherAll([pid]).tolist() == everyone
    assert comm.gatherAll([pid, pid * 2, pid * 3]).tolist() == [x for p in everyone for x in (p, 2 * p, 3 * p)]
    assert comm.gatherAll(np.arange(pid, dtype=np.float64)).tolist() == [float(x) for p in everyone for x in range(p)]  # ragged
    comm.barrier()
    assert comm.sumAll([pid]).tolist() == [tot] and comm.sumAll([8, 3, pid, 2]).tolist() == [8 * P, 3 * P, tot, 2 * P]
    assert comm.maxAll([pid, -pid, 8]).tolist() == [P, -1, 8] and comm.minAll([pid, -pid, 6]).tolist() == [1, -P, 6]
    assert comm.sumAll([0.5 * pid]).tolist() == [0.5 * tot]
    assert comm.scanSum([pid]).tolist() == [pid * (pid + 1) // 2]
    assert comm.scanSum([5, -2 * pid, 3]).tolist() == [pid * 5, sum(range(-2, -2 * pid - 1, -2)), pid * 3]
    # ---- Distributor over the NCCL all-to-all-v: test/MPICommTests.jl:47-65 ----
    d = comm.createDistributor()
    assert d.createFromSends(everyone) == P
    assert d.resolve([pid * j for j in everyone]).tolist() == [pid * j for j in everyone]
    assert d.createFromSends(everyone + everyone) == 2 * P
    got = d.resolve([5 * t + pid for t in range(2 * P)])
    exp = []
    for j in everyone:
        exp += [(pid - 1) * 5 + j, (pid - 1 + P) * 5 + j]
    assert got.tolist() == exp
    eg, ep = d.createFromRecvs([pid + 5 * t for t in range(P)], everyone)
    assert eg.tolist() == [(pid - 1) * 5 + j for j in everyone] and ep.tolist() == everyone

    # ---- test/CSRMatrixMPITests.jl: 1-D Laplacian-like matrix, n=2 rows per rank, k=2 ----
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
    A.fillComplete(m, m)
    Y = jp.DenseMultiVector(m, np.diag(np.arange(1.0, n + 1)))
    X = jp.DenseMultiVector(m, np.full((n, n), 2.0))
    assert A.apply_(Y, X, jp.NO_TRANS, 3.0, 0.5) is Y
    exp = np.diag(np.arange(1.0, n + 1)) * 0.5
    if pid == 1:
        exp[0, :] += 6
    if P > 1 and pid == P:
        exp[n - 1, :] += 6 if P * n > 2 else 0
    if P * n > 2:
        assert np.array_equal(Y.data, exp), (pid, Y.data, exp)
    assert np.array_equal(X.data, np.full((n, n), 2.0))
    Y2 = A * X
    e2 = np.zeros((n, n))
    if pid == 1:
        e2[0, :] = 2
    if pid == P:
        e2[n - 1, :] = 2
    if P * n > 2:
        assert np.array_equal(Y2.data, e2)

    # ---- stencils and a random matrix against the oracle's P-rank simulation ----
    for kind, size, k, exact in (("7pt", 24, 1, True), ("27pt", 16, 4, False), ("5pt", 40, 2, False), ("powerlaw", 30000, 1, False),
                                 ("powerlaw", 8000, 3, False)):
        rm = jp.BlockMap(global_rows(kind, size), comm)
        Aj = synthetic.build_matrix(jp, rm, kind, size, max_len=3000)
        oc, om, OA = oracle_problem(kind, size, P, max_len=3000)
        rp, ci, va = OA.csr(pid)
        assert np.array_equal(Aj.rowOffsets, rp) and np.array_equal(Aj.localIndices1D, ci) and np.array_equal(Aj.values, va)
        assert np.array_equal(Aj.colMap.myGlobalElements(), OA.col_map.my_gids(pid))
        if OA.importer is not None:
            assert_plan_equal(Aj.importer, OA.importer, pid)
        xs = [synthetic.vector_block(om.info(p)["minMyGID"], om.info(p)["numMyElements"], k, seed=1) for p in everyone]
        ys = [synthetic.vector_block(om.info(p)["minMyGID"], om.info(p)["numMyElements"], k, seed=2) for p in everyone]
        OX, OY = O.OMV.from_arrays(om, xs), O.OMV.from_arrays(om, ys)
        OA.apply(OY, OX, O.NO_TRANS, 3.0, 0.5)
        ref = OY.get(pid)
        # component-wise bound |alpha| |A| |x| + |beta| |y0| through the oracle on absolute values (the bound of the
        # single-GPU tests, tests/test_gpu_parity.py)
        _, _, OAabs = oracle_problem(kind, size, P, max_len=3000, abs_vals=True)
        OYb = O.OMV.from_arrays(om, [np.abs(y) for y in ys])
        OAabs.apply(OYb, O.OMV.from_arrays(om, [np.abs(x) for x in xs]), O.NO_TRANS, 3.0, 0.5)
        bound = OYb.get(pid)
        Xj, Yj = jp.DenseMultiVector(rm, xs[rank]), jp.DenseMultiVector(rm, ys[rank])
        for _ in range(3):
            Yj.data = ys[rank]
            Aj.apply_(Yj, Xj, jp.NO_TRANS, 3.0, 0.5)
        got = Yj.data
        err = np.abs(got - ref)
        assert (err <= TOL * bound + 1e-300).all(), (kind, pid, float((err / (bound + 1e-300)).max()))
        if exact:
            assert np.array_equal(got, ref), (kind, pid, np.abs(got - ref).max())
        # host-buffer entry point
        yh = Aj.apply_host(xs[rank], np.array(ys[rank], order="F", copy=True), jp.NO_TRANS, 3.0, 0.5)
        assert np.array_equal(yh, got)
        # beta == 0: the chunked three-stream pipeline of jp_apply_host (H2D chunks | halo + boundary rows | interior chunks | D2H)
        for _ in range(2):
            yh0 = Aj.apply_host(xs[rank], None, jp.NO_TRANS, 3.0, 0.0)
        assert np.array_equal(yh0, Aj.apply(Xj, jp.NO_TRANS, 3.0).data), (kind, pid, "pipelined host path")
        # dot / norm with the cross-rank allreduce
        dref = OX.dot(OY)
        dgot = Xj.dot(Yj)[0]
        dscale = O.OMV.from_arrays(om, [np.abs(x) for x in xs]).dot(O.OMV.from_arrays(om, [np.abs(OY.get(p)) for p in everyone]))
        assert (np.abs(dgot - dref) <= TOL * dscale).all(), (kind, dgot, dref)
        nref = OY.norm(2)
        ngot = Yj.norm(2)[0]
        assert (np.abs(ngot - nref) <= TOL * nref).all()
        # generic doImport into a column-map multivector == the oracle's
        if Aj.importer is not None:
            XC = jp.DenseMultiVector(Aj.colMap, k)
            jp.doImport(Xj, XC, Aj.importer, jp.INSERT)
            OXC = O.OMV(OA.col_map, k)
            O.do_transfer(OX, OXC, OA.importer, False, O.INSERT)
            assert np.array_equal(XC.data, OXC.get(pid)), (kind, pid)
        # device-side fillComplete on every rank: same column map, Import lists, CSR and result
        Ad = synthetic.build_matrix(jp, rm, kind, size, fill=False, max_len=3000)
        Ad.fillComplete(rm, rm, device=True)
        assert Ad._csr_host is None  # nothing was packed on the host
        assert np.array_equal(Ad.colMap.myGlobalElements(), OA.col_map.my_gids(pid))
        assert np.array_equal(Ad.rowOffsets, rp) and np.array_equal(Ad.localIndices1D, ci) and np.array_equal(Ad.values, va)
        if OA.importer is not None:
            assert_plan_equal(Ad.importer, OA.importer, pid)
        Yd = jp.DenseMultiVector(rm, ys[rank])
        Ad.apply_(Yd, Xj, jp.NO_TRANS, 3.0, 0.5)
        assert np.array_equal(Yd.data, got), (kind, pid, "device fillComplete")
        # fused building blocks: apply!+dot and axpby+norm2, against the separate calls
        Yf = jp.DenseMultiVector(rm, ys[rank])
        df = Aj.apply_dot_(Yf, Xj, None, jp.NO_TRANS, 3.0, 0.5)[0]
        assert np.array_equal(Yf.data, got) and (np.abs(df - dref) <= TOL * dscale).all(), (kind, df, dref)
        R1, R2 = jp.DenseMultiVector(rm, ys[rank]), jp.DenseMultiVector(rm, ys[rank])
        R1.axpby_(-0.37, Xj, 1.0)
        n1 = R1.norm(2)[0]
        n2 = R2.axpby_norm2_(-0.37, Xj, 1.0)[0]
        assert np.array_equal(R1.data, R2.data) and (np.abs(n1 - n2) <= TOL * n1).all()
    # ---- one plan, changing multivector widths: k = 2 (arena sized for 4 vectors), k = 6 (the peer-memory arena is regrown
    #      collectively: mappings closed, barrier, old arena retired), k = 1 (fused kernel on the regrown arena), k = 6 again ----
    rm = jp.BlockMap(global_rows("7pt", 20), comm)
    Aj = synthetic.build_matrix(jp, rm, "7pt", 20)
    oc, om, OA = oracle_problem("7pt", 20, P)
    for k in (2, 6, 1, 6):
        xs = [synthetic.vector_block(om.info(p)["minMyGID"], om.info(p)["numMyElements"], k, seed=3) for p in everyone]
        OX, OY = O.OMV.from_arrays(om, xs), O.OMV(om, k)
        OA.apply(OY, OX, O.NO_TRANS, 1.0, 0.0)
        Yk = Aj.apply(jp.DenseMultiVector(rm, xs[rank]))
        ref = OY.get(pid)
        assert np.abs(Yk.data - ref).max() <= 1e-12 * 12.0, (pid, k)
        if k == 1:
            assert np.array_equal(Yk.data, ref)
    # ---- distributed TRANS (Tpetra semantics, see oracle/jp_oracle.cpp applyMatrix) ----
    for kind, size, k in (("7pt", 12, 2), ("powerlaw", 6000, 1), ("27pt", 8, 3)):
        rm = jp.BlockMap(global_rows(kind, size), comm)
        Aj = synthetic.build_matrix(jp, rm, kind, size, max_len=500)
        oc, om, OA = oracle_problem(kind, size, P, max_len=500)
        xs = [synthetic.vector_block(om.info(p)["minMyGID"], om.info(p)["numMyElements"], k, seed=1) for p in everyone]
        ys = [synthetic.vector_block(om.info(p)["minMyGID"], om.info(p)["numMyElements"], k, seed=2) for p in everyone]
        OX, OY = O.OMV.from_arrays(om, xs), O.OMV.from_arrays(om, ys)
        OA.apply(OY, OX, O.TRANS, 3.0, 0.5)
        Xj, Yj = jp.DenseMultiVector(rm, xs[rank]), jp.DenseMultiVector(rm, ys[rank])
        Aj.apply_(Yj, Xj, jp.TRANS, 3.0, 0.5)
        ref = OY.get(pid)
        # component-wise bound |beta| |y0| + |alpha| |A|^T |x| through the oracle on absolute values
        _, _, OAabs = oracle_problem(kind, size, P, max_len=500, abs_vals=True)
        OYb = O.OMV.from_arrays(om, [np.abs(y) for y in ys])
        OAabs.apply(OYb, O.OMV.from_arrays(om, [np.abs(x) for x in xs]), O.TRANS, 3.0, 0.5)
        err = np.abs(Yj.data - ref)
        assert (err <= TOL * OYb.get(pid) + 1e-300).all(), (kind, pid, float((err / (OYb.get(pid) + 1e-300)).max()))
    # ---- a column map with unused local columns (permutes) AND remotes: the generic import + single-launch path ----
    Ng = 40 * P + 3
    rng = np.random.default_rng(11)
    cols_of = {g: sorted(set(rng.integers(1, Ng + 1, size=4).tolist()) - {2, 7}) or [1] for g in range(1, Ng + 1)}  # GIDs 2, 7 are never columns
    vals_of = {g: rng.standard_normal(len(cols_of[g])) for g in range(1, Ng + 1)}
    mg = jp.BlockMap(Ng, comm)
    Ag = jp.CSRMatrix(mg, [len(cols_of[int(g)]) for g in mg.myGlobalElements()], jp.STATIC_PROFILE)
    for g in mg.myGlobalElements():
        Ag.insertGlobalValues(int(g), cols_of[int(g)], vals_of[int(g)])
    Ag.fillComplete(mg, mg)
    ocg = O.OComm(P)
    omg = O.OMap.uniform(ocg, Ng)
    OAg = O.OMatrix(omg, [[len(cols_of[int(g)]) for g in omg.my_gids(p)] for p in everyone])
    for p in everyone:
        for g in omg.my_gids(p):
            OAg.insert(p, int(g), cols_of[int(g)], vals_of[int(g)])
    OAg.fill_complete(omg, omg)
    for k in (1, 3):
        xs = [synthetic.vector_block(omg.info(p)["minMyGID"], omg.info(p)["numMyElements"], k, seed=1) for p in everyone]
        ys = [synthetic.vector_block(omg.info(p)["minMyGID"], omg.info(p)["numMyElements"], k, seed=2) for p in everyone]
        OX, OY = O.OMV.from_arrays(omg, xs), O.OMV.from_arrays(omg, ys)
        OAg.apply(OY, OX, O.NO_TRANS, 3.0, 0.5)
        Xj, Yj = jp.DenseMultiVector(mg, xs[rank]), jp.DenseMultiVector(mg, ys[rank])
        Ag.apply_(Yj, Xj, jp.NO_TRANS, 3.0, 0.5)
        triples = [(g, c, v) for g in range(1, Ng + 1) for c, v in zip(cols_of[g], vals_of[g])]
        lo, hi = omg.info(pid)["minMyGID"] - 1, omg.info(pid)["maxMyGID"]
        assert (np.abs(Yj.data - OY.get(pid)) <= TOL * dense_abs_bound(triples, Ng, xs, ys, False, lo, hi) + 1e-300).all(), (pid, k)
        OY2 = O.OMV.from_arrays(omg, ys)
        OAg.apply(OY2, OX, O.TRANS, 3.0, 0.5)
        Yt = jp.DenseMultiVector(mg, ys[rank])
        Ag.apply_(Yt, Xj, jp.TRANS, 3.0, 0.5)
        assert (np.abs(Yt.data - OY2.get(pid)) <= TOL * dense_abs_bound(triples, Ng, xs, ys, True, lo, hi) + 1e-300).all(), (pid, k, "trans")
    # ---- CG on the distributed 7-point Laplacian: host scalars (fused / separate building blocks) and device scalars ----
    rm = jp.BlockMap(20 ** 3, comm)
    Ac = synthetic.build_matrix(jp, rm, "7pt", 20)
    xs_true = synthetic.vector_block(rm.minMyGID(), rm.numMyElements(), 1, seed=7)
    Bc = Ac.apply(jp.DenseMultiVector(rm, xs_true))
    for fused in (True, False):
        Xc = jp.DenseMultiVector(rm, 1)
        its, rel = jp.cg(Ac, Bc, Xc, tol=1e-12, maxiter=400, fused=fused)
        assert rel <= 1e-12 and np.abs(Xc.data - xs_true).max() <= 1e-9, (fused, its, rel)
    Xc = jp.DenseMultiVector(rm, 1)
    its, rel = jp.cg_device(Ac, Bc, Xc, tol=1e-12, maxiter=400, check_every=5)
    assert rel <= 1e-12 and np.abs(Xc.data - xs_true).max() <= 1e-9, ("device scalars", its, rel)
    # ---- rangeMap != rowMap: an operator assembled on an overlapping row map ----
    from fill_cases import overlapping_rows_problem
    oc, ouni, orow, OAe, per_rank, dense = overlapping_rows_problem(P, 40)
    rows_e, rp_e, cols_e, vals_e = per_rank[rank]
    uni = jp.BlockMap(P * 40, comm)
    rowmap = jp.BlockMap(-1, rows_e, comm)
    Ae = jp.CSRMatrix(rowmap, np.diff(rp_e), jp.STATIC_PROFILE)
    Ae.insertGlobalValuesBatch(rows_e, rp_e, cols_e, vals_e)
    Ae.fillComplete(uni, uni)
    assert Ae.getExporter() is not None
    assert_plan_equal(Ae.exporter, OAe.exporter, pid)
    for k in (1, 3):
        xs = [synthetic.vector_block(ouni.info(p)["minMyGID"], ouni.info(p)["numMyElements"], k, seed=1) for p in everyone]
        ys = [synthetic.vector_block(ouni.info(p)["minMyGID"], ouni.info(p)["numMyElements"], k, seed=2) for p in everyone]
        for mode, omode in ((jp.NO_TRANS, O.NO_TRANS), (jp.TRANS, O.TRANS)):
            OX, OY = O.OMV.from_arrays(ouni, xs), O.OMV.from_arrays(ouni, ys)
            OAe.apply(OY, OX, omode, 3.0, 0.5)
            Xe, Ye = jp.DenseMultiVector(uni, xs[rank]), jp.DenseMultiVector(uni, ys[rank])
            for _ in range(2):
                Ye.data = ys[rank]
                Ae.apply_(Ye, Xe, mode, 3.0, 0.5)
            ref = OY.get(pid)
            triples = [(int(g), int(cols_p[i]), vals_p[i]) for rows_p, rp_p, cols_p, vals_p in per_rank for r, g in enumerate(rows_p)
                       for i in range(rp_p[r], rp_p[r + 1])]
            lo, hi = ouni.info(pid)["minMyGID"] - 1, ouni.info(pid)["maxMyGID"]
            bound = dense_abs_bound(triples, P * 40, xs, ys, mode == jp.TRANS, lo, hi)
            assert (np.abs(Ye.data - ref) <= TOL * bound + 1e-300).all(), ("exporter", pid, k, int(mode))
    # commReduce on a replicated map (test/DenseMultiVectorTests.jl:96-99): 18 doubles go through the peer-memory allreduce
    # kernel, 150 doubles through ncclAllReduce
    rep = jp.BlockMap(6, 6, comm)
    v = jp.DenseMultiVector(rep, (10.0 ** pid) * np.ones((6, 3)))
    v.commReduce()
    assert np.array_equal(v.data, sum(10.0 ** i for i in everyone) * np.ones((6, 3)))
    v = jp.DenseMultiVector(rep, (2.0 ** pid) * np.ones((6, 25)))
    v.commReduce()
    assert np.array_equal(v.data, float(sum(2 ** i for i in everyone)) * np.ones((6, 25)))
    # a multivector wider than the mapped result block's fast path is still reduced correctly (k = 70 dot products)
    wide = jp.BlockMap(P * 50, comm)
    xw = synthetic.vector_block(wide.minMyGID(), wide.numMyElements(), 70, seed=4)
    Xw = jp.DenseMultiVector(wide, xw)
    dw = Xw.dot(Xw)[0]
    allx = [synthetic.vector_block(1 + 50 * (p - 1), 50, 70, seed=4) for p in everyone]
    dref = sum((a * a).sum(axis=0) for a in allx)
    assert np.abs(dw - dref).max() <= 1e-12 * dref.max()
    comm.barrier()
    dist.barrier()
    if rank == 0:
        print(f"multi-GPU parity ok on {world} ranks", flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
