# TevoB200.jl - thin `ccall` shim that puts libtevo.so (include/tevo.h) behind TimeEvoMPS.jl's own call sites.
#
# NOT EXECUTED IN THIS REPOSITORY'S CI: neither Julia nor ITensors.jl exists in the build image (SURVEY.md 0.3), so
# this file is reviewed, not run.  It is deliberately mechanical: every function is one `ccall` plus the index-order
# marshalling ITensors needs.  The identical symbols are exercised from Python (timeevomps.jl_b200/_lib.py) by the
# parity tests, which is what pins their behaviour.
#
# Usage inside the reference package (replaces the bodies of src/bondgate.jl:46-63 and the loop body of
# src/tdvp.jl:56-88; everything else - BondOperator, gates, exp, time_evo_gates, tebd!'s scheduler, the callbacks -
# stays as it is):
#
#     using TevoB200
#     dpsi = TevoB200.upload(psi)                       # once per tebd!/tdvp! call
#     spec = TevoB200.apply_gate!(dpsi, G; maxdim=..., cutoff=...)      # instead of apply_gate!(psi, G; ...)
#     TevoB200.download!(psi, dpsi)                     # before returning / for user callbacks
module TevoB200

using ITensors

const LIB = get(ENV, "TEVO_LIB", joinpath(@__DIR__, "..", "libtevo.so"))

struct TevoTrunc
    has_maxdim::Cint
    maxdim::Cint
    mindim::Cint
    has_cutoff::Cint
    cutoff::Cdouble
    use_absolute_cutoff::Cint
    do_rel_cutoff::Cint
end

struct TevoSpec
    nkeep::Cint
    nfull::Cint
    truncerr::Cdouble
    entropy::Cdouble
    sum_all::Cdouble
    sum_kept::Cdouble
end

struct TevoKrylov
    hermitian::Cint
    tol::Cdouble
    krylovdim::Cint
    maxiter::Cint
end

struct TevoKrylovInfo
    converged::Cint
    numops::Cint
    numiter::Cint
end

const TEVO_ENOTCONVERGED = 4

mutable struct Ctx
    h::Ptr{Cvoid}
end

function Ctx(device::Integer=0)
    r = Ref{Ptr{Cvoid}}(C_NULL)
    st = ccall((:tevo_ctx_create, LIB), Cint, (Cint, Ref{Ptr{Cvoid}}), device, r)
    st == 0 || error(unsafe_string(ccall((:tevo_last_error, LIB), Cstring, (Ptr{Cvoid},), C_NULL)))
    c = Ctx(r[])
    finalizer(x -> ccall((:tevo_ctx_destroy, LIB), Cint, (Ptr{Cvoid},), x.h), c)
    return c
end

check(ctx::Ctx, st) = st == 0 ? nothing :
    throw(unsafe_string(ccall((:tevo_last_error, LIB), Cstring, (Ptr{Cvoid},), ctx.h)))

"Device-resident image of an ITensors MPS; keeps the site indices so that download! can rebuild ITensors."
mutable struct DeviceMPS
    ctx::Ctx
    h::Ptr{Cvoid}
    sites::Vector{<:Index}
    linktags::Vector{TagSet}
end

trunc_from(; maxdim=nothing, mindim=1, cutoff=nothing, use_absolute_cutoff=false, absoluteCutoff=false,
           doRelCutoff=true, kwargs...) =
    TevoTrunc(maxdim === nothing ? 0 : 1, maxdim === nothing ? 0 : maxdim, mindim,
              cutoff === nothing ? 0 : 1, cutoff === nothing ? 0.0 : cutoff,
              (use_absolute_cutoff || absoluteCutoff) ? 1 : 0, doRelCutoff ? 1 : 0)

"Permute site tensor j of `psi` to (left link, site, right link) and return a dense ComplexF64 array."
function site_array(psi::MPS, j::Int)
    N = length(psi)
    s = siteind(psi, j)
    l = j > 1 ? commonind(psi[j - 1], psi[j]) : nothing
    r = j < N ? commonind(psi[j], psi[j + 1]) : nothing
    is = filter(!isnothing, (l, s, r))
    A = Array{ComplexF64}(array(permute(psi[j], is...)))
    chil = l === nothing ? 1 : dim(l)
    chir = r === nothing ? 1 : dim(r)
    return reshape(A, chil, dim(s), chir)
end

function upload(psi::MPS; ctx::Ctx=Ctx())
    N = length(psi)
    d = dim(siteind(psi, 1))
    r = Ref{Ptr{Cvoid}}(C_NULL)
    check(ctx, ccall((:tevo_mps_create, LIB), Cint, (Ptr{Cvoid}, Cint, Cint, Ref{Ptr{Cvoid}}), ctx.h, N, d, r))
    dpsi = DeviceMPS(ctx, r[], [siteind(psi, j) for j in 1:N], [tags(linkind(psi, j)) for j in 1:N-1])
    finalizer(x -> ccall((:tevo_mps_free, LIB), Cint, (Ptr{Cvoid},), x.h), dpsi)
    for j in 1:N
        A = site_array(psi, j)
        GC.@preserve A check(ctx, ccall((:tevo_mps_upload_site, LIB), Cint,
                                        (Ptr{Cvoid}, Cint, Cint, Cint, Ptr{ComplexF64}),
                                        dpsi.h, j - 1, size(A, 1), size(A, 3), A))
    end
    # ITensors leftlim/rightlim are 1-based "last left-orthogonal"/"first right-orthogonal": subtract 1
    check(ctx, ccall((:tevo_mps_set_limits, LIB), Cint, (Ptr{Cvoid}, Cint, Cint),
                     dpsi.h, ITensors.leftlim(psi) - 1, ITensors.rightlim(psi) - 1))
    return dpsi
end

"Copy the device state back into `psi` (new link indices are created with the old tags, as replacebond! does)."
function download!(psi::MPS, dpsi::DeviceMPS)
    N = length(psi)
    dims = Vector{Cint}(undef, N - 1)
    check(dpsi.ctx, ccall((:tevo_mps_bonddims, LIB), Cint, (Ptr{Cvoid}, Ptr{Cint}), dpsi.h, dims))
    links = [Index(dims[b], dpsi.linktags[b]) for b in 1:N-1]
    for j in 1:N
        chil = j > 1 ? dims[j - 1] : 1
        chir = j < N ? dims[j] : 1
        A = Array{ComplexF64}(undef, chil, dim(dpsi.sites[j]), chir)
        GC.@preserve A check(dpsi.ctx, ccall((:tevo_mps_download_site, LIB), Cint,
                                             (Ptr{Cvoid}, Cint, Ptr{ComplexF64}, Csize_t), dpsi.h, j - 1, A, length(A)))
        is = filter(!isnothing, (j > 1 ? links[j - 1] : nothing, dpsi.sites[j], j < N ? links[j] : nothing))
        psi[j] = ITensor(reshape(A, map(dim, is)...), is...)
    end
    ll, rl = Ref{Cint}(0), Ref{Cint}(0)
    check(dpsi.ctx, ccall((:tevo_mps_get_limits, LIB), Cint, (Ptr{Cvoid}, Ref{Cint}, Ref{Cint}), dpsi.h, ll, rl))
    ITensors.setleftlim!(psi, ll[] + 1)
    ITensors.setrightlim!(psi, rl[] + 1)
    return psi
end

"Row-major d^2 x d^2 matrix G[(s1',s2'),(s1,s2)] of a BondGate ITensor (primed = out, src/bondgate.jl:24)."
function gate_matrix(G::ITensor, s1::Index, s2::Index)
    A = Array{ComplexF64}(array(permute(G, s2', s1', s2, s1)))   # column-major (s2',s1',s2,s1) == row-major ((s1',s2'),(s1,s2))
    return vec(A)
end

"apply_gate!(psi, G; kwargs...) (src/bondgate.jl:46-53) on the device state; returns ITensors.Spectrum."
function apply_gate!(dpsi::DeviceMPS, G::ITensor, b::Int; ortho="left", normalize=true, kwargs...)
    g = gate_matrix(G, dpsi.sites[b], dpsi.sites[b + 1])
    tr = Ref(trunc_from(; kwargs...))
    spec = Ref(TevoSpec(0, 0, 0.0, 0.0, 0.0, 0.0))
    cap = 4096
    eigs = Vector{Cdouble}(undef, cap)
    GC.@preserve g eigs check(dpsi.ctx, ccall((:tevo_apply_gate, LIB), Cint,
        (Ptr{Cvoid}, Cint, Ptr{ComplexF64}, Ref{TevoTrunc}, Cint, Cint, Ref{TevoSpec}, Ptr{Cdouble}, Cint),
        dpsi.h, b - 1, g, tr, ortho == "left" ? 1 : 0, normalize ? 1 : 0, spec, eigs, cap))
    return ITensors.Spectrum(eigs[1:spec[].nkeep], spec[].truncerr)
end

"apply_gates!(psi, Gs; reversed, cb, kwargs...) (src/bondgate.jl:55-63): one device call per Trotter layer."
function apply_gates!(dpsi::DeviceMPS, Gs::Vector{<:Tuple{ITensor,Int}}; reversed=false, cb=nothing, normalize=true, kwargs...)
    n = length(Gs)
    bonds = Cint[b - 1 for (_, b) in Gs]
    gs = reduce(vcat, [gate_matrix(G, dpsi.sites[b], dpsi.sites[b + 1]) for (G, b) in Gs])
    tr = Ref(trunc_from(; kwargs...))
    specs = Vector{TevoSpec}(undef, n)
    stride = 4096
    eigs = cb === nothing ? Cdouble[] : Vector{Cdouble}(undef, n * stride)
    GC.@preserve bonds gs specs eigs check(dpsi.ctx, ccall((:tevo_apply_layer, LIB), Cint,
        (Ptr{Cvoid}, Cint, Ptr{Cint}, Ptr{ComplexF64}, Cint, Ref{TevoTrunc}, Cint, Cint, Ptr{TevoSpec}, Ptr{Cdouble}, Cint,
         Cint, Ptr{ComplexF64}, Ptr{ComplexF64}),
        dpsi.h, n, bonds, gs, reversed ? 1 : 0, tr, normalize ? 1 : 0, 0, specs,
        cb === nothing ? C_NULL : pointer(eigs), cb === nothing ? 0 : stride, 0, C_NULL, C_NULL))
    if cb !== nothing
        for g in (reversed ? (n:-1:1) : (1:n))
            k = specs[g].nkeep
            cb(dpsi; bond=Gs[g][2], spec=ITensors.Spectrum(eigs[(g-1)*stride+1:(g-1)*stride+k], specs[g].truncerr))
        end
    end
end

"<O> on sites b and b+1 from theta = psi[b]*psi[b+1] (src/callback.jl:65-73); ops: d x d matrices O[s',s]."
function measure_bond(dpsi::DeviceMPS, b::Int, ops::Vector{Matrix{ComplexF64}})
    flat = reduce(vcat, [vec(permutedims(o)) for o in ops])      # row-major
    out = Vector{ComplexF64}(undef, 2 * length(ops))
    GC.@preserve flat out check(dpsi.ctx, ccall((:tevo_measure_bond, LIB), Cint,
        (Ptr{Cvoid}, Cint, Cint, Ptr{ComplexF64}, Ptr{ComplexF64}), dpsi.h, b - 1, length(ops), flat, out))
    return reshape(out, 2, length(ops))
end

# ---- state maintenance and observables on the device ---------------------------------------------------
"orthogonalize!(psi, j) (src/bondgate.jl:48, src/tdvp.jl:51)."
orthogonalize!(dpsi::DeviceMPS, j::Int) =
    check(dpsi.ctx, ccall((:tevo_mps_orthogonalize, LIB), Cint, (Ptr{Cvoid}, Cint), dpsi.h, j - 1))

"reorthogonalize!(psi) (src/itensor.jl:28-33, called at src/tebd.jl:174)."
reorthogonalize!(dpsi::DeviceMPS) = check(dpsi.ctx, ccall((:tevo_mps_reorthogonalize, LIB), Cint, (Ptr{Cvoid},), dpsi.h))

"psi[i] /= norm(psi[i]) (src/tdvp.jl:84-87)."
normalize_site!(dpsi::DeviceMPS, i::Int) =
    check(dpsi.ctx, ccall((:tevo_mps_normalize_site, LIB), Cint, (Ptr{Cvoid}, Cint), dpsi.h, i - 1))

"deepcopy(psi) on the device."
function Base.copy(dpsi::DeviceMPS)
    r = Ref{Ptr{Cvoid}}(C_NULL)
    check(dpsi.ctx, ccall((:tevo_mps_copy, LIB), Cint, (Ptr{Cvoid}, Ref{Ptr{Cvoid}}), dpsi.h, r))
    c = DeviceMPS(dpsi.ctx, r[], dpsi.sites, dpsi.linktags)
    finalizer(x -> ccall((:tevo_mps_free, LIB), Cint, (Ptr{Cvoid},), x.h), c)
    return c
end

"inner(a, b) = <a|b> (test/test_tebd.jl:16,28,46)."
function inner(a::DeviceMPS, b::DeviceMPS)
    out = Ref{ComplexF64}(0)
    check(a.ctx, ccall((:tevo_mps_inner, LIB), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ref{ComplexF64}), a.h, b.h, out))
    return out[]
end

"measure!(psi, op, j) (src/testutils.jl:55-58): moves the centre to j and returns <O_j>; O[s',s] as a d x d matrix."
function expect(dpsi::DeviceMPS, O::Matrix{ComplexF64}, j::Int)
    flat = vec(permutedims(O))                                   # row-major
    out = Ref{ComplexF64}(0)
    GC.@preserve flat check(dpsi.ctx, ccall((:tevo_mps_expect_site, LIB), Cint,
        (Ptr{Cvoid}, Cint, Ptr{ComplexF64}, Ref{ComplexF64}), dpsi.h, j - 1, flat, out))
    return out[]
end

"One term of measure(H::GateList, psi) (src/bondgate.jl:71-79): <h_b> for the bond operator h (d^2 x d^2, as gate_matrix)."
function gate_expect(dpsi::DeviceMPS, h::ITensor, b::Int)
    g = gate_matrix(h, dpsi.sites[b], dpsi.sites[b + 1])
    out = Ref{ComplexF64}(0)
    GC.@preserve g check(dpsi.ctx, ccall((:tevo_gate_expect, LIB), Cint,
        (Ptr{Cvoid}, Cint, Ptr{ComplexF64}, Ref{ComplexF64}), dpsi.h, b - 1, g, out))
    return out[]
end

# ---- layer-parallel mode (NOT the reference's algorithm, see DESIGN.md section 8) ------------------------
"Bring the state to the canonical B form with Schmidt values on every bond; the gates of a layer become independent."
to_bform!(dpsi::DeviceMPS) = check(dpsi.ctx, ccall((:tevo_mps_to_bform, LIB), Cint, (Ptr{Cvoid},), dpsi.h))

"All gates of one even/odd layer at once (Hastings' update); same arguments as apply_gates! without `reversed`."
function apply_gates_parallel!(dpsi::DeviceMPS, Gs::Vector{<:Tuple{ITensor,Int}}; normalize=true, kwargs...)
    n = length(Gs)
    bonds = Cint[b - 1 for (_, b) in Gs]
    gs = reduce(vcat, [gate_matrix(G, dpsi.sites[b], dpsi.sites[b + 1]) for (G, b) in Gs])
    tr = Ref(trunc_from(; kwargs...))
    specs = Vector{TevoSpec}(undef, n)
    GC.@preserve bonds gs specs check(dpsi.ctx, ccall((:tevo_apply_layer_bform, LIB), Cint,
        (Ptr{Cvoid}, Cint, Ptr{Cint}, Ptr{ComplexF64}, Ref{TevoTrunc}, Cint, Ptr{TevoSpec}, Ptr{Cdouble}, Cint),
        dpsi.h, n, bonds, gs, tr, normalize ? 1 : 0, specs, C_NULL, 0))
    return specs
end

"<O_j> for several one-site operators in the B form (no centre move)."
function bform_expect(dpsi::DeviceMPS, ops::Vector{Matrix{ComplexF64}}, j::Int)
    flat = reduce(vcat, [vec(permutedims(o)) for o in ops])
    out = Vector{ComplexF64}(undef, length(ops))
    GC.@preserve flat out check(dpsi.ctx, ccall((:tevo_bform_expect_site, LIB), Cint,
        (Ptr{Cvoid}, Cint, Cint, Ptr{ComplexF64}, Ptr{ComplexF64}), dpsi.h, j - 1, length(ops), flat, out))
    return out
end

"Schmidt values of bond j (between sites j and j+1) in the B form."
function schmidt_values(dpsi::DeviceMPS, j::Int)
    buf = Vector{Cdouble}(undef, 65536)
    nout = Ref{Cint}(0)
    GC.@preserve buf check(dpsi.ctx, ccall((:tevo_mps_get_lambda, LIB), Cint,
        (Ptr{Cvoid}, Cint, Ptr{Cdouble}, Cint, Ref{Cint}), dpsi.h, j, buf, length(buf), nout))   # C: bond left of 0-based site j
    return buf[1:nout[]]
end

synchronize(ctx::Ctx) = check(ctx, ccall((:tevo_ctx_synchronize, LIB), Cint, (Ptr{Cvoid},), ctx.h))

# ---- TDVP (src/tdvp.jl:52-88) -------------------------------------------------------------------------
mutable struct DeviceProjMPO
    ctx::Ctx
    hmpo::Ptr{Cvoid}
    h::Ptr{Cvoid}
end

function DeviceProjMPO(H::MPO, dpsi::DeviceMPS)
    N = length(H)
    d = dim(dpsi.sites[1])
    r = Ref{Ptr{Cvoid}}(C_NULL)
    check(dpsi.ctx, ccall((:tevo_mpo_create, LIB), Cint, (Ptr{Cvoid}, Cint, Cint, Ref{Ptr{Cvoid}}), dpsi.ctx.h, N, d, r))
    hm = r[]
    for j in 1:N
        s = dpsi.sites[j]
        l = j > 1 ? commonind(H[j - 1], H[j]) : nothing
        rr = j < N ? commonind(H[j], H[j + 1]) : nothing
        is = filter(!isnothing, (l, s', s, rr))
        W = Array{ComplexF64}(array(permute(H[j], is...)))
        wl = l === nothing ? 1 : dim(l)
        wr = rr === nothing ? 1 : dim(rr)
        GC.@preserve W check(dpsi.ctx, ccall((:tevo_mpo_upload_site, LIB), Cint,
                                             (Ptr{Cvoid}, Cint, Cint, Cint, Ptr{ComplexF64}), hm, j - 1, wl, wr, W))
    end
    e = Ref{Ptr{Cvoid}}(C_NULL)
    check(dpsi.ctx, ccall((:tevo_env_create, LIB), Cint, (Ptr{Cvoid}, Ref{Ptr{Cvoid}}), hm, e))
    PH = DeviceProjMPO(dpsi.ctx, hm, e[])
    finalizer(PH) do x
        ccall((:tevo_env_free, LIB), Cint, (Ptr{Cvoid},), x.h)
        ccall((:tevo_mpo_free, LIB), Cint, (Ptr{Cvoid},), x.hmpo)
    end
    return PH
end

"position!(PH, psi, pos) with PH.nsite = nsite (src/tdvp.jl:53,58-59,79-80); twosite!/onesite! call it themselves."
position!(PH::DeviceProjMPO, dpsi::DeviceMPS, pos::Int; nsite::Int=2) =
    check(dpsi.ctx, ccall((:tevo_env_position, LIB), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Cint, Cint), PH.h, dpsi.h, pos - 1, nsite))

"inner(psi, H, phi) (test/test_tdvp.jl:29-33) on the device."
function inner(a::DeviceMPS, PH::DeviceProjMPO, b::DeviceMPS)
    out = Ref{ComplexF64}(0)
    check(a.ctx, ccall((:tevo_mps_inner_mpo, LIB), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ref{ComplexF64}),
                       a.h, PH.hmpo, b.h, out))
    return out[]
end

"Forward two-site step of the sweep (src/tdvp.jl:58-64); throws like :63 when exponentiate does not converge."
function twosite!(PH::DeviceProjMPO, dpsi::DeviceMPS, b::Int, t::Number; hermitian=true, exp_tol=1e-14, krylovdim=30,
                  ortho="left", normalize=true, kwargs...)
    kp = Ref(TevoKrylov(hermitian ? 1 : 0, exp_tol, krylovdim, 100))
    tr = Ref(trunc_from(; kwargs...))
    spec = Ref(TevoSpec(0, 0, 0.0, 0.0, 0.0, 0.0))
    info = Ref(TevoKrylovInfo(0, 0, 0))
    eigs = Vector{Cdouble}(undef, 4096)
    tc = ComplexF64(t)
    st = GC.@preserve eigs ccall((:tevo_tdvp_twosite, LIB), Cint,
        (Ptr{Cvoid}, Ptr{Cvoid}, Cint, ComplexF64, Ref{TevoKrylov}, Ref{TevoTrunc}, Cint, Cint, Ref{TevoSpec}, Ptr{Cdouble},
         Cint, Ref{TevoKrylovInfo}),
        PH.h, dpsi.h, b - 1, tc, kp, tr, ortho == "left" ? 1 : 0, normalize ? 1 : 0, spec, eigs, 4096, info)
    st == TEVO_ENOTCONVERGED && throw("exponentiate did not converge")
    check(dpsi.ctx, st)
    return ITensors.Spectrum(eigs[1:spec[].nkeep], spec[].truncerr)
end

"One-site backward step (src/tdvp.jl:77-83)."
function onesite!(PH::DeviceProjMPO, dpsi::DeviceMPS, i::Int, t::Number; hermitian=true, exp_tol=1e-14, krylovdim=30,
                  maxiter=100)
    kp = Ref(TevoKrylov(hermitian ? 1 : 0, exp_tol, krylovdim, maxiter))
    info = Ref(TevoKrylovInfo(0, 0, 0))
    st = ccall((:tevo_tdvp_onesite, LIB), Cint,
               (Ptr{Cvoid}, Ptr{Cvoid}, Cint, ComplexF64, Ref{TevoKrylov}, Ref{TevoKrylovInfo}),
               PH.h, dpsi.h, i - 1, ComplexF64(t), kp, info)
    st == TEVO_ENOTCONVERGED && throw("exponentiate did not converge")
    check(dpsi.ctx, st)
end

# ---- drop-in drivers: tebd!(psi::MPS, ...) / tdvp!(psi::MPS, H::MPO, ...) behind the reference signatures ------------
# The schedule (time_evo_gates, bunching, direction alternation, callback arguments) is TimeEvoMPS.jl's own code path
# restated line by line from src/tebd.jl:91-179 and src/tdvp.jl:37-96; only the arithmetic call sites are replaced:
#   apply_gates!(psi, U; ...)            -> TevoB200.apply_gates!(dpsi, U; ...)          (src/bondgate.jl:55-63)
#   position! / exponentiate / replacebond! -> TevoB200.twosite! / onesite!              (src/tdvp.jl:58-64,77-83)
# The state lives on the device for the whole call: upload once, download once at the end.  Callbacks keep the
# reference protocol apply!(cb, psi; t, bond, spec, sweepdir, sweepend, alg); what they may read from `psi` decides
# how much is copied back before each invocation (state_needed below).
import TimeEvoMPS
import TimeEvoMPS: TEBDalg, TEBD2, TEBD4, TDVP2, BondGate, GateList, BondOperator, gates, time_evo_gates, bond,
                   TEvoCallback, NoTEvoCallback, LocalMeasurementCallback, SpecCallback, apply!, checkdone!, callback_dt

"What a callback reads from `psi`: :nothing (spec only), :bond (psi[bond], psi[bond+1] at sweep-end gates, as
LocalMeasurementCallback does at src/callback.jl:89) or :all (unknown user callback: the whole state, every gate)."
state_needed(::NoTEvoCallback) = :nothing
state_needed(::SpecCallback) = :nothing
state_needed(::LocalMeasurementCallback) = :bond
state_needed(::TEvoCallback) = :all

"Copy sites `js` of the device state back into `psi` (link indices are renewed consistently for all of them)."
function download_sites!(psi::MPS, dpsi::DeviceMPS, js)
    N = length(psi)
    dims = Vector{Cint}(undef, N - 1)
    check(dpsi.ctx, ccall((:tevo_mps_bonddims, LIB), Cint, (Ptr{Cvoid}, Ptr{Cint}), dpsi.h, dims))
    # links touching the requested sites get fresh indices; a site outside `js` that shares one of them keeps a stale
    # tensor, which is why :bond callbacks may only contract psi[bond]*psi[bond+1] (exactly what callback.jl:89 does)
    newlink = Dict{Int,Index}()
    for j in js, b in (j - 1, j)
        1 <= b <= N - 1 && !haskey(newlink, b) && (newlink[b] = Index(dims[b], dpsi.linktags[b]))
    end
    for j in js
        chil = j > 1 ? dims[j - 1] : 1
        chir = j < N ? dims[j] : 1
        A = Array{ComplexF64}(undef, chil, dim(dpsi.sites[j]), chir)
        GC.@preserve A check(dpsi.ctx, ccall((:tevo_mps_download_site, LIB), Cint,
                                             (Ptr{Cvoid}, Cint, Ptr{ComplexF64}, Csize_t), dpsi.h, j - 1, A, length(A)))
        is = filter(!isnothing, (j > 1 ? newlink[j - 1] : nothing, dpsi.sites[j], j < N ? newlink[j] : nothing))
        psi[j] = ITensor(reshape(A, map(dim, is)...), is...)
    end
    return psi
end

function sync_for_callback!(psi::MPS, dpsi::DeviceMPS, cb, b::Int, sweepend::Bool)
    need = state_needed(cb)
    need === :all && return download!(psi, dpsi)
    need === :bond && sweepend && return download_sites!(psi, dpsi, (b, b + 1))
    return psi
end

"tebd!(psi, H, dt, tf, alg; kwargs...) with the arithmetic on the B200 (src/tebd.jl:89-179, same keywords)."
tebd!(psi::MPS, H::BondOperator, args...; kwargs...) = tebd!(psi, gates(H), args...; kwargs...)
function tebd!(psi::MPS, H::GateList, dt::Number, tf::Number, alg::TEBDalg=TEBD2(); ctx::Ctx=Ctx(), kwargs...)
    nsteps = Int(tf / dt)                                                        # :95 (InexactError as upstream)
    ishermitian = get(kwargs, :ishermitian, false)                               # :100
    cb = get(kwargs, :callback, NoTEvoCallback())                                # :102
    orthogonalize_step = get(kwargs, :orthogonalize, 0)                          # :110
    dtm = callback_dt(cb)
    if dtm > 0                                                                   # :113-120
        floor(Int64, dtm / dt) != dtm / dt &&
            throw("Measurement time step $dtm incommensurate with time-evolution time step $dt")
        nbunch = gcd(floor(Int64, dtm / dt), nsteps)
    else
        nbunch = nsteps
    end
    orthogonalize_step > 0 && (nbunch = gcd(nbunch, orthogonalize_step))         # :121
    Ustart, Us, Uend = time_evo_gates(dt, H, alg; ishermitian=ishermitian)       # :123
    length(Ustart) == 0 && length(Uend) == 0 && (nbunch += 1)                    # :125
    trunc_kw = filter(kv -> kv[1] in (:maxdim, :mindim, :cutoff, :use_absolute_cutoff, :normalize), kwargs)
    dpsi = upload(psi; ctx=ctx)
    pairs_of(U) = Tuple{ITensor,Int}[(g.G, bond(g)) for g in U]
    function cb_func(dt, step, rev, se)                                          # :103-109
        cb isa NoTEvoCallback && return nothing
        return (d; bond, spec) -> begin
            sync_for_callback!(psi, d, cb, bond, se)
            apply!(cb, psi; t=dt * step, bond=bond, spec=spec, sweepdir=rev ? "left" : "right", sweepend=se, alg=alg)
        end
    end
    step = 0
    rev = false
    while step < nsteps                                                          # :132
        for U in Ustart                                                          # :133-136
            apply_gates!(dpsi, pairs_of(U); reversed=rev, trunc_kw...)
            rev = !rev
        end
        for i in 1:nbunch-1                                                      # :138-153
            for U in Us
                apply_gates!(dpsi, pairs_of(U); reversed=rev, cb=cb_func(dt, step, rev, false), trunc_kw...)
                rev = !rev
            end
            step += 1
            checkdone!(cb, psi) && break
        end
        for (i, U) in enumerate(Uend)                                            # :157-160
            apply_gates!(dpsi, pairs_of(U); reversed=rev, cb=cb_func(dt, step + 1, rev, i >= length(Uend) - 1), trunc_kw...)
            rev = !rev
        end
        length(Uend) > 0 && (step += 1)                                          # :163
        (orthogonalize_step > 0 && step % orthogonalize_step == 0) && reorthogonalize!(dpsi)   # :174
        checkdone!(cb, psi) && break                                             # :176
    end
    download!(psi, dpsi)
    return psi
end

"tdvp!(psi, H::MPO, dt, tf; kwargs...) with the arithmetic on the B200 (src/tdvp.jl:37-96, same keywords)."
function tdvp!(psi::MPS, H::MPO, dt::Number, tf::Number; ctx::Ctx=Ctx(), kwargs...)
    nsteps = Int(tf / dt)                                                        # :38
    cb = get(kwargs, :callback, NoTEvoCallback())
    hermitian = get(kwargs, :hermitian, true)
    exp_tol = get(kwargs, :exp_tol, 1e-14)
    krylovdim = get(kwargs, :krylovdim, 30)
    maxiter = get(kwargs, :maxiter, 100)
    normalize = get(kwargs, :normalize, true)                                    # :39-44
    τ = 1im * dt
    imag(τ) == 0 && (τ = real(τ))                                                # :47-48
    N = length(psi)
    trunc_kw = filter(kv -> kv[1] in (:maxdim, :mindim, :cutoff, :use_absolute_cutoff, :absoluteCutoff), kwargs)
    dpsi = upload(psi; ctx=ctx)
    orthogonalize!(dpsi, 1)                                                      # :51
    PH = DeviceProjMPO(H, dpsi)                                                  # :52
    position!(PH, dpsi, 1)                                                       # :53
    for s in 1:nsteps                                                            # :54
        for (b, ha) in sweepnext(N)                                              # :56
            spec = twosite!(PH, dpsi, b, -τ / 2; hermitian=hermitian, exp_tol=exp_tol, krylovdim=krylovdim,
                            ortho=ha == 1 ? "left" : "right", normalize=normalize, trunc_kw...)      # :58-64
            if !(cb isa NoTEvoCallback)
                sync_for_callback!(psi, dpsi, cb, b, ha == 2)
                apply!(cb, psi; t=s * dt, bond=b, sweepend=ha == 2, sweepdir=ha == 1 ? "right" : "left", spec=spec,
                       alg=TDVP2())                                              # :67-72
            end
            i = ha == 1 ? b + 1 : b                                              # :77
            if 1 < i < N && !(dt isa Complex)
                onesite!(PH, dpsi, i, τ / 2; hermitian=hermitian, exp_tol=exp_tol, krylovdim=krylovdim, maxiter=maxiter)  # :78-83
            elseif i == 1 && dt isa Complex
                normalize_site!(dpsi, i)                                         # :84-87
            end
        end
        checkdone!(cb) && break                                                  # :94
    end
    download!(psi, dpsi)
    return nothing
end

end # module
