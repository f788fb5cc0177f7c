# JuliaPetraB200.jl -- the reference-side binding a JuliaPetra maintainer adds to route the hot path
# (CSRMatrix apply!, Import halo, MultiVector dot/norm) through libjpetra_b200.so.
#
# NOT EXECUTED IN THIS REPOSITORY: the build image has no Julia (SURVEY.md Appendix C).  The Python host mirror
# (juliapetra.jl_b200/*.py) drives the identical C ABI in the identical call order and is what the tests run;
# this file shows, entry point by entry point, what each `ccall` replaces in the reference.  Everything that is
# not listed here (BlockMap, Directory, fillComplete, Import/Export construction, ...) stays as it is in Julia.
module JuliaPetraB200

using JuliaPetra
import JuliaPetra: barrier, broadcastAll, gatherAll, sumAll, maxAll, minAll, scanSum, myPid, numProc,
                   createDistributor, apply!, localApply, dot, norm, scale!, getLocalArray, numVectors, getMap,
                   commReduce, doTransfer, fillComplete

const LIB = get(ENV, "JPETRA_B200_LIB", "libjpetra_b200.so")
const Handle = Int64

# ---- error convention: status codes -> the reference's exceptions (src/Utils.jl:37-51) -----------------------
function check(rc::Int32)
    rc == 0 && return
    buf = Vector{UInt8}(undef, 2048)
    ccall((:jp_last_error, LIB), Int32, (Ptr{UInt8}, Int64), buf, length(buf))
    msg = unsafe_string(pointer(buf))
    rc == 1 && throw(InvalidArgumentError(msg))
    rc == 2 && throw(InvalidStateError(msg))
    error("jpetra_b200 backend failure ($rc): $msg")
end

init(device::Integer = parse(Int, get(ENV, "LOCAL_RANK", "0"))) = check(ccall((:jp_init, LIB), Int32, (Int32,), device))

# ---- Comm: NCCLComm replaces MPIComm (src/MPIComm.jl:1-91) behind src/Comm.jl:12-57 ----------------------------
struct NCCLComm{GID <: Integer, PID <: Integer, LID <: Integer} <: Comm{GID, PID, LID}
    handle::Handle
    pid::PID       # 1-based, src/MPIComm.jl:81-83
    nproc::PID
end

"""
    NCCLComm{GID,PID,LID}(uid::Vector{UInt8}, nranks, rank0)

`uid` is the 128-byte id rank 0 obtains from `uniqueId()` and the launcher distributes (environment, file, ...).
"""
function NCCLComm{GID, PID, LID}(uid::Vector{UInt8}, nranks::Integer, rank0::Integer) where {GID, PID, LID}
    h = Ref{Handle}(0)
    check(ccall((:jp_comm_create_nccl, LIB), Int32, (Ptr{UInt8}, Int32, Int32, Ref{Handle}), uid, nranks, rank0, h))
    NCCLComm{GID, PID, LID}(h[], PID(rank0 + 1), PID(nranks))
end

function uniqueId()
    uid = Vector{UInt8}(undef, 128)
    check(ccall((:jp_comm_unique_id, LIB), Int32, (Ptr{UInt8},), uid))
    uid
end

myPid(c::NCCLComm) = c.pid
numProc(c::NCCLComm) = c.nproc
barrier(c::NCCLComm) = check(ccall((:jp_comm_barrier, LIB), Int32, (Handle,), c.handle))

dtypecode(::Type{Float64}) = Int32(0)
dtypecode(::Type{Int64}) = Int32(1)

function allreduce(c::NCCLComm, op::Integer, v::AbstractArray{T}) where T <: Union{Float64, Int64}
    out = similar(v)
    check(ccall((:jp_comm_allreduce, LIB), Int32, (Handle, Int32, Int32, Ptr{T}, Ptr{T}, Int64),
                c.handle, op, dtypecode(T), v, out, length(v)))
    out
end
# other integer widths go through Int64, Bool through Int64 like the UInt8 detour of src/MPIComm.jl:65-75
allreduce(c::NCCLComm, op::Integer, v::AbstractArray{T}) where T <: Integer = Array{T}(allreduce(c, op, Array{Int64}(v)))

sumAll(c::NCCLComm, v::AbstractArray) = allreduce(c, 0, v)     # src/MPIComm.jl:57-59
maxAll(c::NCCLComm, v::AbstractArray) = allreduce(c, 1, v)     # src/MPIComm.jl:61-67
minAll(c::NCCLComm, v::AbstractArray) = allreduce(c, 2, v)     # src/MPIComm.jl:69-75

function scanSum(c::NCCLComm, v::AbstractArray{T}) where T <: Union{Float64, Int64}   # src/MPIComm.jl:77-79
    out = similar(v)
    check(ccall((:jp_comm_scan_sum, LIB), Int32, (Handle, Int32, Ptr{T}, Ptr{T}, Int64), c.handle, dtypecode(T), v, out, length(v)))
    out
end

function gatherAll(c::NCCLComm, v::AbstractArray{T}) where T <: Union{Float64, Int64}  # src/MPIComm.jl:52-55
    counts = zeros(Int64, c.nproc)
    check(ccall((:jp_comm_gather_all, LIB), Int32, (Handle, Int32, Ptr{Int64}, Int64, Ptr{Int64}, Ptr{Int64}),
                c.handle, 1, Int64[length(v)], 1, ones(Int64, c.nproc), counts))
    out = Vector{T}(undef, sum(counts))
    check(ccall((:jp_comm_gather_all, LIB), Int32, (Handle, Int32, Ptr{T}, Int64, Ptr{Int64}, Ptr{T}),
                c.handle, dtypecode(T), v, length(v), counts, out))
    out
end

function broadcastAll(c::NCCLComm, v::AbstractArray{T}, root::Integer) where T <: Union{Float64, Int64}  # :46-50
    out = copy(v)
    check(ccall((:jp_comm_broadcast, LIB), Int32, (Handle, Int32, Ptr{T}, Int64, Int32), c.handle, dtypecode(T), out, length(out), root))
    out
end

# The Distributor plan construction (createFromSends / createFromRecvs, src/MPIDistributor.jl:69-299) stays in
# Julia; its point-to-point traffic becomes one call:
function alltoallv(c::NCCLComm, send::Vector{T}, sendcounts::Vector{Int64}, recvcounts::Vector{Int64}) where T
    recv = Vector{T}(undef, sum(recvcounts))
    check(ccall((:jp_comm_alltoallv, LIB), Int32, (Handle, Int32, Ptr{T}, Ptr{Int64}, Ptr{T}, Ptr{Int64}),
                c.handle, sizeof(T), send, sendcounts, recv, recvcounts))
    recv
end

# ---- Distributor: NCCLDistributor replaces MPIDistributor (src/MPIDistributor.jl) behind src/Distributor.jl:10-41 ----
# Same plan record and the same field values (they are bit-exact parity targets); what changes is how the plan is
# learnt (one gatherAll of a P-vector of message lengths instead of Reduce+Scatter, ANY_SOURCE receives and
# ready-sends, :205-251) and how host records move (one all-to-all-v instead of 2 handshakes + 4 barriers, :487-590).
mutable struct NCCLDistributor{GID <: Integer, PID <: Integer, LID <: Integer} <: Distributor{GID, PID, LID}
    comm::NCCLComm{GID, PID, LID}
    lengths_to::Vector{Int64}; procs_to::Vector{PID}; indices_to::Vector{Int64}; starts_to::Vector{Int64}
    lengths_from::Vector{Int64}; procs_from::Vector{PID}; starts_from::Vector{Int64}
    numSends::Int64; numRecvs::Int64; selfMsg::Int64; totalRecvLength::Int64
    posted::Union{Vector, Nothing}
    planReverse::Union{NCCLDistributor{GID, PID, LID}, Nothing}
end
NCCLDistributor(comm::NCCLComm{GID, PID, LID}) where {GID, PID, LID} =
    NCCLDistributor{GID, PID, LID}(comm, [], [], [], [], [], [], [], 0, 0, 0, 0, nothing, nothing)
createDistributor(comm::NCCLComm) = NCCLDistributor(comm)

function JuliaPetra.createFromSends(dist::NCCLDistributor{GID, PID, LID}, exportPIDs::AbstractArray{PID}) where {GID, PID, LID}
    pid, nProcs = myPid(dist.comm), numProc(dist.comm)
    counts = zeros(Int64, nProcs)                      # createSendStructure, src/MPIDistributor.jl:69-193
    for p in exportPIDs; p >= 1 && (counts[p] += 1); end
    grouped = all(exportPIDs[i] >= exportPIDs[i-1] for i in 2:length(exportPIDs))
    dist.selfMsg = counts[pid] != 0
    dist.procs_to = PID[p for p in 1:nProcs if counts[p] > 0]
    dist.lengths_to = counts[dist.procs_to]
    ndead = count(<(1), exportPIDs)
    if grouped
        dist.indices_to = Int64[]
        dist.starts_to = (ndead + 1) .+ vcat(0, cumsum(dist.lengths_to)[1:end-1])
    else
        active = findall(>=(1), exportPIDs)
        dist.indices_to = active[sortperm(exportPIDs[active])]
        dist.starts_to = 1 .+ vcat(0, cumsum(dist.lengths_to)[1:end-1])
    end
    dist.numSends = length(dist.procs_to) - dist.selfMsg
    sendLen = zeros(Int64, nProcs); sendLen[dist.procs_to] = dist.lengths_to   # computeRecvs, :196-273
    M = reshape(gatherAll(dist.comm, sendLen), nProcs, nProcs)                # M[dst, src]
    col = M[pid, :]
    dist.procs_from = PID[q for q in 1:nProcs if col[q] > 0]
    dist.lengths_from = col[dist.procs_from]
    dist.starts_from = 1 .+ vcat(0, cumsum(dist.lengths_from)[1:end-1])
    dist.totalRecvLength = sum(dist.lengths_from)
    dist.numRecvs = length(dist.procs_from) - dist.selfMsg
    dist.planReverse = nothing
    dist.totalRecvLength
end

function JuliaPetra.createFromRecvs(dist::NCCLDistributor{GID, PID, LID}, remoteGIDs::AbstractArray{GID}, remotePIDs::AbstractArray{PID}) where {GID, PID, LID}
    length(remoteGIDs) == length(remotePIDs) || throw(InvalidArgumentError("remote lists must be the same length"))
    tmp = NCCLDistributor(dist.comm)                   # computeSends, :275-299
    JuliaPetra.createFromSends(tmp, copy(remotePIDs))
    exportObjs = JuliaPetra.resolve(tmp, [(Int64(g), Int64(myPid(dist.comm))) for g in remoteGIDs])
    exportGIDs = GID[o[1] for o in exportObjs]; exportPIDs = PID[o[2] for o in exportObjs]
    JuliaPetra.createFromSends(dist, exportPIDs)
    exportGIDs, exportPIDs
end

function JuliaPetra.resolvePosts(dist::NCCLDistributor, exportObjs::AbstractVector{T}) where T   # :451-564
    nProcs = numProc(dist.comm)
    send = isempty(dist.indices_to) ? exportObjs[1:sum(dist.lengths_to)] : exportObjs[dist.indices_to]
    sc = zeros(Int64, nProcs); sc[dist.procs_to] = dist.lengths_to
    rc = zeros(Int64, nProcs); rc[dist.procs_from] = dist.lengths_from
    dist.posted = alltoallv(dist.comm, Vector{T}(send), sc, rc)       # arrives in procs_from order (:582-589)
    nothing
end
function JuliaPetra.resolveWaits(dist::NCCLDistributor)                                          # :568-590
    dist.posted === nothing && throw(InvalidStateError("Cannot resolve waits when no posts have been made"))
    r = dist.posted; dist.posted = nothing; r
end

# ---- DenseMultiVector storage on the device (src/DenseMultiVector.jl:6-60) -------------------------------------------
# The reference documents createViews / createViewsNonConst / releaseViews as the hooks "to fetch data from ... GPU
# into host memory" (src/DistObject.jl:260-284).  A device multivector keeps a handle and a lazily synchronised
# host mirror; getLocalArray downloads when the device copy is newer.
mutable struct DeviceMultiVector{Data, GID, PID, LID} <: MultiVector{Data, GID, PID, LID}
    handle::Handle
    numVectors::LID
    map::BlockMap{GID, PID, LID}
    host::Matrix{Data}
    hostStale::Bool
end

function DeviceMultiVector(map::BlockMap{GID, PID, LID}, data::Matrix{Float64}) where {GID, PID, LID}
    h = Ref{Handle}(0)
    check(ccall((:jp_mv_create, LIB), Int32, (Int32, Int32, Ref{Handle}), size(data, 1), size(data, 2), h))
    check(ccall((:jp_mv_upload, LIB), Int32, (Handle, Ptr{Float64}), h[], data))
    mv = DeviceMultiVector{Float64, GID, PID, LID}(h[], LID(size(data, 2)), map, data, false)
    finalizer(m -> ccall((:jp_mv_destroy, LIB), Int32, (Handle,), m.handle), mv)
    mv
end

numVectors(mv::DeviceMultiVector) = mv.numVectors
getMap(mv::DeviceMultiVector) = mv.map
function getLocalArray(mv::DeviceMultiVector)
    if mv.hostStale
        check(ccall((:jp_mv_download, LIB), Int32, (Handle, Ptr{Float64}), mv.handle, mv.host))
        mv.hostStale = false
    end
    mv.host
end

commhandle(mv::DeviceMultiVector) = JuliaPetra.getComm(mv.map).handle

function dot(x::DeviceMultiVector{Float64}, y::DeviceMultiVector{Float64})   # src/MultiVector.jl:84-108
    out = Matrix{Float64}(undef, 1, numVectors(x))
    check(ccall((:jp_mv_dot, LIB), Int32, (Handle, Handle, Handle, Ptr{Float64}), commhandle(x), x.handle, y.handle, out))
    out
end

function norm(x::DeviceMultiVector{Float64}, n::Real)                         # src/MultiVector.jl:115-146
    out = Matrix{Float64}(undef, 1, numVectors(x))
    if n == 2
        check(ccall((:jp_mv_norm2, LIB), Int32, (Handle, Handle, Ptr{Float64}), commhandle(x), x.handle, out))
    else
        check(ccall((:jp_mv_normp, LIB), Int32, (Handle, Handle, Float64, Ptr{Float64}), commhandle(x), x.handle, n, out))
    end
    out
end

function scale!(x::DeviceMultiVector{Float64}, alpha::Number)                 # src/MultiVector.jl:62-65
    check(ccall((:jp_mv_scale, LIB), Int32, (Handle, Float64), x.handle, alpha)); x.hostStale = true; x
end
function Base.fill!(x::DeviceMultiVector{Float64}, v)                         # src/MultiVector.jl:52-55
    check(ccall((:jp_mv_fill, LIB), Int32, (Handle, Float64), x.handle, v)); x.hostStale = true; x
end
function commReduce(x::DeviceMultiVector{Float64})                            # src/MultiVector.jl:154-161
    check(ccall((:jp_mv_comm_reduce, LIB), Int32, (Handle, Handle), commhandle(x), x.handle)); x.hostStale = true; x
end
"y = a*x + b*y -- the axpy of a CG loop, what `@. y = a*x + b*y` lowers to (src/MultiVector.jl:497-502)"
function axpby!(y::DeviceMultiVector{Float64}, a, x::DeviceMultiVector{Float64}, b)
    check(ccall((:jp_mv_axpby, LIB), Int32, (Handle, Float64, Handle, Float64), y.handle, a, x.handle, b)); y.hostStale = true; y
end

"""
    axpbyNorm2!(y, a, x, b)

`y = a*x + b*y` and `norm(y, 2)` of the result in one pass (`jp_mv_axpby_norm2`): the residual update of a CG
iteration fused with its norm (broadcast copyto!, src/MultiVector.jl:497-502, + norm, :122-133).
"""
function axpbyNorm2!(y::DeviceMultiVector{Float64}, a, x::DeviceMultiVector{Float64}, b)
    out = Matrix{Float64}(undef, 1, numVectors(y))
    check(ccall((:jp_mv_axpby_norm2, LIB), Int32, (Handle, Handle, Float64, Handle, Float64, Ptr{Float64}),
                commhandle(y), y.handle, a, x.handle, b, out))
    y.hostStale = true
    out
end

# ---- device CSR + plan, created once at the end of fillComplete (src/CSRMatrix.jl:706-747) ----------------------------
const DEVICE_CSR = IdDict{Any, Handle}()     # CSRMatrix => jp csr handle
const DEVICE_PLAN = IdDict{Any, Handle}()    # Import     => jp plan handle

function upload!(A::CSRMatrix{Float64, GID, PID, LID}) where {GID, PID, LID}
    g = A.myGraph
    ncols = numMyElements(getColMap(A))
    imp = getImporter(A)
    nlocal = imp === nothing ? ncols : numSameIDs(imp)
    h = Ref{Handle}(0)
    check(ccall((:jp_csr_create, LIB), Int32,
                (Int32, Int32, Int32, Int64, Ptr{Int32}, Ptr{Int32}, Ptr{Float64}, Int32, Ref{Handle}),
                getLocalNumRows(A), ncols, nlocal, length(A.localMatrix.values),
                Vector{Int32}(g.rowOffsets), Vector{Int32}(g.localIndices1D), A.localMatrix.values, 1, h))
    DEVICE_CSR[A] = h[]
    if imp !== nothing
        d = distributor(imp)                    # the plan fields of src/MPIDistributor.jl:12-65
        p = Ref{Handle}(0)
        check(ccall((:jp_plan_create, LIB), Int32,
                    (Handle, Int32, Int32, Int32, Ptr{Int32}, Ptr{Int32}, Int32, Ptr{Int32}, Int32, Ptr{Int32}, Int32,
                     Ptr{Int32}, Ptr{Int64}, Int32, Ptr{Int32}, Ptr{Int64}, Int32, Int32, Ref{Handle}),
                    JuliaPetra.getComm(A).handle, numMyElements(sourceMap(imp)), numMyElements(targetMap(imp)), numSameIDs(imp),
                    Vector{Int32}(permuteToLIDs(imp)), Vector{Int32}(permuteFromLIDs(imp)), length(permuteToLIDs(imp)),
                    Vector{Int32}(remoteLIDs(imp)), length(remoteLIDs(imp)),
                    Vector{Int32}(exportLIDs(imp)), length(exportLIDs(imp)),
                    Vector{Int32}(d.procs_to), Vector{Int64}(d.lengths_to), length(d.procs_to),
                    Vector{Int32}(d.procs_from), Vector{Int64}(d.lengths_from), length(d.procs_from), 1, p))
        DEVICE_PLAN[imp] = p[]
    end
    A
end

"""
    fillCompleteDevice!(A, domainMap, rangeMap)

`fillComplete` with the O(nnz) part on the device (`jp_csr_fill_complete`): the entries inserted so far go over as they
are (global column ids, insertion order); the column map (src/CSRGraphInternalMethods.jl:852-982), the GID -> LID remap
(:636-712), the per-row sort + duplicate merge (src/CSRMatrix.jl:549-606) and the CSR packing (:403-484) run as
kernels.  The host then only builds the column-map `BlockMap` from the returned GIDs and the `Import` from it, as
`fillComplete` does today.  Needs a contiguous domain map (`BlockMap(n, comm)` / `BlockMap(n, nMy, comm)`).
"""
function fillCompleteDevice!(A::CSRMatrix{Float64, GID, PID, LID}, domainMap::BlockMap{GID, PID, LID},
                             rangeMap::BlockMap{GID, PID, LID}) where {GID, PID, LID}
    g = A.myGraph
    nrows = getLocalNumRows(A)
    counts = g.numRowEntries                                # entries actually inserted per row
    ptr0 = Vector{Int64}(undef, nrows + 1); ptr0[1] = 0
    for r in 1:nrows ptr0[r+1] = ptr0[r] + counts[r] end
    gcols = Vector{Int64}(undef, ptr0[end]); vals = Vector{Float64}(undef, ptr0[end])
    for r in 1:nrows                                        # pack the STATIC_PROFILE storage (src/CSRMatrix.jl:617-665)
        s = g.rowOffsets[r]
        gcols[ptr0[r]+1:ptr0[r+1]] = g.globalIndices1D[s:s+counts[r]-1]
        vals[ptr0[r]+1:ptr0[r+1]] = A.values[s:s+counts[r]-1]
    end
    h = Ref{Handle}(0); info = Vector{Int64}(undef, 4)
    rc = ccall((:jp_csr_fill_complete, LIB), Int32,
               (Int32, Ptr{Int64}, Ptr{Int64}, Ptr{Float64}, Int64, Int64, Int64, Int32, Ref{Handle}, Ptr{Int64}),
               nrows, ptr0, gcols, vals, minAllGID(domainMap), maxAllGID(domainMap),
               numMyElements(domainMap) > 0 ? minMyGID(domainMap) : minAllGID(domainMap), numMyElements(domainMap), h, info)
    comm = JuliaPetra.getComm(A)
    err = rc == 1                                           # a column GID nobody owns (src/CSRGraphInternalMethods.jl:936-939)
    rc > 1 && check(rc)
    maxAll(comm, err ? 1 : 0) != 0 && error("makeColMap reports an error on at least one process")
    colGIDs = Vector{GID}(undef, info[1])
    check(ccall((:jp_csr_col_map, LIB), Int32, (Handle, Ptr{Int64}), h[], colGIDs))
    colMap = BlockMap(numGlobalElements(domainMap), colGIDs, comm)      # src/CSRGraphInternalMethods.jl:981
    # ... set A's colMap / domainMap / rangeMap / importer / exporter exactly as fillComplete does (makeImportExport,
    # src/CSRGraphInternalMethods.jl:351-366), then remember the device matrix instead of calling upload!:
    DEVICE_CSR[A] = h[]
    colMap
end

"""
    applyDot!(Y, A, X, W, mode, alpha, beta)

`apply!(Y, A, X, mode, alpha, beta)` followed by `dot(W, Y)` as one call (`jp_apply_dot`); for one vector without an
importer the SpMV kernel accumulates the dot product itself (the p'Ap of CG with W = X).
"""
function applyDot!(Y::DeviceMultiVector{Float64, GID, PID, LID}, A::CSRMatrix{Float64, GID, PID, LID},
                   X::DeviceMultiVector{Float64, GID, PID, LID}, W::DeviceMultiVector{Float64, GID, PID, LID},
                   mode::TransposeMode, alpha::Float64, beta::Float64) where {GID, PID, LID}
    isFillActive(A) && throw(InvalidStateError("Cannot call apply(...) until fillComplete(...)"))
    haskey(DEVICE_CSR, A) || upload!(A)
    imp = getImporter(A)
    out = Matrix{Float64}(undef, 1, numVectors(Y))
    check(ccall((:jp_apply_dot, LIB), Int32, (Handle, Handle, Handle, Handle, Handle, Int32, Float64, Float64, Handle, Ptr{Float64}),
                JuliaPetra.getComm(A).handle, Y.handle, DEVICE_CSR[A], X.handle, imp === nothing ? 0 : DEVICE_PLAN[imp], Int32(mode),
                alpha, beta, W.handle, out))
    Y.hostStale = true
    out
end

# ---- Operator: apply! (src/RowMatrix.jl:261-357) on device multivectors ---------------------------------------------------
function apply!(Y::DeviceMultiVector{Float64, GID, PID, LID}, A::CSRMatrix{Float64, GID, PID, LID},
                X::DeviceMultiVector{Float64, GID, PID, LID}, mode::TransposeMode, alpha::Float64, beta::Float64) where {GID, PID, LID}
    isFillActive(A) && throw(InvalidStateError("Cannot call apply(...) until fillComplete(...)"))
    haskey(DEVICE_CSR, A) || upload!(A)
    imp = getImporter(A)
    comm = JuliaPetra.getComm(A)
    YIsReplicated = !distributedGlobal(Y)
    if alpha != 0 && YIsReplicated && myPid(comm) != 1
        beta = 0.0                                   # src/RowMatrix.jl:283-287
    end
    ex = getExporter(A)
    if ex === nothing
        check(ccall((:jp_apply, LIB), Int32, (Handle, Handle, Handle, Handle, Handle, Int32, Float64, Float64),
                    comm.handle, Y.handle, DEVICE_CSR[A], X.handle, imp === nothing ? 0 : DEVICE_PLAN[imp], Int32(mode), alpha, beta))
    else
        # rangeMap != rowMap (src/RowMatrix.jl:298-308, 324-330): DEVICE_PLAN[ex] is made from the Export's lists with the
        # same jp_plan_create call upload! uses for the Import
        check(ccall((:jp_apply_export, LIB), Int32, (Handle, Handle, Handle, Handle, Handle, Handle, Int32, Float64, Float64),
                    comm.handle, Y.handle, DEVICE_CSR[A], X.handle, imp === nothing ? 0 : DEVICE_PLAN[imp], DEVICE_PLAN[ex], Int32(mode),
                    alpha, beta))
    end
    Y.hostStale = true
    alpha != 0 && YIsReplicated && commReduce(Y)     # src/RowMatrix.jl:353-355
    Y
end

# apply! on plain host DenseMultiVectors: one call, H2D + kernels + D2H inside (the e2e path of bench.py)
function apply!(Y::DenseMultiVector{Float64, GID, PID, LID}, A::CSRMatrix{Float64, GID, PID, LID},
                X::DenseMultiVector{Float64, GID, PID, LID}, mode::TransposeMode, alpha::Float64, beta::Float64) where {GID, PID, LID}
    isFillActive(A) && throw(InvalidStateError("Cannot call apply(...) until fillComplete(...)"))
    haskey(DEVICE_CSR, A) || upload!(A)
    imp = getImporter(A)
    comm = JuliaPetra.getComm(A)
    GC.@preserve X Y check(ccall((:jp_apply_host, LIB), Int32,
                (Handle, Handle, Handle, Ptr{Float64}, Ptr{Float64}, Int32, Int32, Float64, Float64),
                comm.handle, DEVICE_CSR[A], imp === nothing ? 0 : DEVICE_PLAN[imp], X.data, Y.data, numVectors(X), Int32(mode), alpha, beta))
    Y
end

# doTransfer for device multivectors (src/DistObject.jl:200-258): copyAndPermute / packAndPrepare / resolve /
# unpackAndCombine all happen inside jp_transfer
function doImportDevice(src::DeviceMultiVector, tgt::DeviceMultiVector, imp::Import, cm::CombineMode; reverse = false)
    check(ccall((:jp_transfer, LIB), Int32, (Handle, Handle, Handle, Int32, Int32), DEVICE_PLAN[imp], src.handle, tgt.handle, Int32(cm), reverse))
    tgt.hostStale = true
    nothing
end

end # module
