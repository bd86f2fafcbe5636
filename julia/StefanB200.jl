# StefanB200.jl -- thin `ccall` layer over libstefan_b200.so that keeps the reference's
# call surface (ArjunNarayanan/StefanProblem: InteriorPenalty, LocalDG, DG1D,
# FiniteDifference), so that the reference's driver scripts switch to the B200 path by
# changing their `include`s.  1:1 with the Python twin in stefanproblem_b200/ (which is
# what the test-suite drives: there is no Julia runtime in the build image, so THIS FILE
# HAS NOT BEEN EXECUTED).  See INTEGRATION.md for the mapping and an example.
#
# Conventions kept from the reference: mutate-in-place `assemble_*!(sysmatrix, ...)`, caller
# owns `sysmatrix` / `sysrhs`, `sparse_operator(sysmatrix, mesh, dofspernode)` and
# `rhs_vector(sysrhs, mesh, dofspernode)` are the consumers.  Difference: `SystemMatrix`
# records WHICH passes were requested; `sparse_operator` returns a matrix-free device
# operator `A` with `A * x` (CuArray{Float64} in, CuArray out; CUDA.jl supplies the device
# pointers) and `coo(A)` for the triplets the reference would have appended.
module StefanB200

using CUDA  # device arrays / pointers / streams only

const lib = get(ENV, "STEFAN_B200_LIB", joinpath(@__DIR__, "..", "stefanproblem_b200", "lib", "libstefan_b200.so"))

struct StefanError <: Exception
    status::Int
    msg::String
end
function check(status::Integer)
    status == 0 && return
    msg = unsafe_string(ccall((:stefan_last_error, lib), Cstring, ()))
    throw(StefanError(status, msg))
end

# ------------------------------------------------------------------ mesh / basis stand-ins
struct LagrangeTensorProductBasis
    dim::Int
    order::Int
end
number_of_basis_functions(b::LagrangeTensorProductBasis) = (b.order + 1)^b.dim

"Uncut Cartesian DG mesh: what DGMesh + CutMesh + MergedMesh! give when no cell is cut."
struct UniformMesh
    x0::Vector{Float64}
    widths::Vector{Float64}
    nelements::Vector{Int}
    basis::LagrangeTensorProductBasis
    cellsign::Vector{Int8}          # +1 / -1 per cell, cells y-fastest
end
UniformMesh(x0, widths, nelements, basis) =
    UniformMesh(x0, widths, nelements, basis, ones(Int8, prod(nelements)))
element_size(m::UniformMesh) = m.widths ./ m.nelements
number_of_cells(m::UniformMesh) = prod(m.nelements)

struct Quadratures
    numqp::Int
end
CellQuadratures(mesh, numqp) = Quadratures(numqp)
FaceQuadratures(mesh, numqp) = Quadratures(numqp)
InterfaceQuadratures(mesh, numqp) = Quadratures(numqp)

mutable struct SystemMatrix
    kind::Symbol
    spec::Dict{Symbol,Any}
    terms::Int
    SystemMatrix() = new(:none, Dict{Symbol,Any}(), 0)
end
mutable struct SystemRHS
    kind::Symbol
    spec::Dict{Symbol,Any}
    SystemRHS() = new(:none, Dict{Symbol,Any}())
end
function record!(s::SystemMatrix, kind::Symbol, terms::Int; kw...)
    s.kind == :none || s.kind == kind || error("SystemMatrix already holds a $(s.kind) operator")
    s.kind = kind
    s.terms |= terms
    for (k, v) in kw
        haskey(s.spec, k) && s.spec[k] != v && error("inconsistent $k: $(s.spec[k]) vs $v")
        s.spec[k] = v
    end
end

# ---------------------------------------------------------------------- InteriorPenalty
module InteriorPenalty
import ..record!, ..SystemMatrix, ..SystemRHS
const VOLUME, IE_FLUX, IE_PENALTY, IF_FLUX, IF_PENALTY, B_FLUX, B_PENALTY, ROBIN = 1, 2, 4, 8, 16, 32, 64, 128

# IP_gradient_operator.jl:29-57
assemble_gradient_operator!(s, basis, cellquads, k1, k2, mesh) =
    record!(s, :ip, VOLUME; order = basis.order, numqp = cellquads.numqp, k1 = k1, k2 = k2)
# IP_flux_operator.jl:202-256
assemble_interelement_flux_operator!(s, basis, facequads, k1, k2, mesh) =
    record!(s, :ip, IE_FLUX; order = basis.order, numqp = facequads.numqp, k1 = k1, k2 = k2)
# IP_penalty_operator.jl:134-181
assemble_interelement_penalty_operator!(s, basis, facequads, penalty, mesh) =
    record!(s, :ip, IE_PENALTY; order = basis.order, numqp = facequads.numqp, penalty = penalty)
# IP_flux_operator.jl:296-324 / IP_penalty_operator.jl:213-237: cut cells only, none on an uncut mesh
assemble_interface_flux_operator!(s, basis, interfacequads, k1, k2, mesh) = nothing
assemble_interface_penalty_operator!(s, basis, interfacequads, penalty, mesh) = nothing
# IP_flux_operator.jl:393-444
assemble_boundary_flux_operator!(s, basis, facequads, k1, k2, mesh) =
    record!(s, :ip, B_FLUX; order = basis.order, numqp = facequads.numqp, k1 = k1, k2 = k2)
# IP_penalty_operator.jl:286-329
assemble_boundary_penalty_operator!(s, basis, facequads, penalty, mesh) =
    record!(s, :ip, B_PENALTY; order = basis.order, numqp = facequads.numqp, penalty = penalty)
# mesh-aligned interface: the reference's face-level functions on unlike-phase mesh faces
assemble_aligned_interface_flux_operator!(s, basis, facequads, k1, k2, mesh) =
    record!(s, :ip, IF_FLUX; order = basis.order, numqp = facequads.numqp, k1 = k1, k2 = k2)
assemble_aligned_interface_penalty_operator!(s, basis, facequads, penalty, mesh) =
    record!(s, :ip, IF_PENALTY; order = basis.order, numqp = facequads.numqp, penalty = penalty)
# IP_robin_interface_operator.jl:76-99
assemble_robin_operator!(s, basis, quads, lambda, mesh) = record!(s, :ip, ROBIN; lambda = lambda)

# IP_assembly.jl:1-80, same order of passes
function assemble_interior_penalty_linear_system!(s, solverbasis, cellquads, facequads, interfacequads,
                                                  k1, k2, penalty, mesh)
    assemble_gradient_operator!(s, solverbasis, cellquads, k1, k2, mesh)
    assemble_interelement_flux_operator!(s, solverbasis, facequads, k1, k2, mesh)
    assemble_interelement_penalty_operator!(s, solverbasis, facequads, penalty, mesh)
    assemble_interface_flux_operator!(s, solverbasis, interfacequads, k1, k2, mesh)
    assemble_interface_penalty_operator!(s, solverbasis, interfacequads, penalty, mesh)
    assemble_boundary_flux_operator!(s, solverbasis, facequads, k1, k2, mesh)
    assemble_boundary_penalty_operator!(s, solverbasis, facequads, penalty, mesh)
end
end # module InteriorPenalty

# ------------------------------------------------------------------------------ LocalDG
module LocalDG
import ..record!, ..SystemMatrix, ..SystemRHS
# LDG_assembly.jl:1-91: divergence + mass + gradient cell operators, inter-element scalar and
# vector fluxes with the switch V0 (passed positionally as `beta`, as the reference does),
# boundary vector flux; on an uncut mesh the reference's interface passes find no cut cell
function assemble_LDG_linear_system!(s, basis, cellquads, facequads, interfacequads, k1, k2, interiorpenalty,
                                     interfacepenalty, negboundarypenalty, posboundarypenalty, V0, mesh)
    record!(s, :ldg, 63; order = basis.order, numqp = cellquads.numqp, k1 = k1, k2 = k2,
            interiorpenalty = interiorpenalty, negboundarypenalty = negboundarypenalty,
            posboundarypenalty = posboundarypenalty, V0 = (V0[1], V0[2]))
end
# mesh-aligned interface: the face-level functions (LDG_scalar_flux_operator.jl:25-107,
# LDG_vector_flux_operator.jl:47-154) on the faces between uncut cells of unlike phase
assemble_aligned_interface_scalar_flux_operator!(s, basis, facequads, V0, mesh) =
    record!(s, :ldg, 0; V0 = (V0[1], V0[2]))
assemble_aligned_interface_vector_flux_operator!(s, basis, facequads, k1, k2, alpha, beta, mesh) =
    record!(s, :ldg, 0; k1 = k1, k2 = k2, interfacepenalty = alpha, V0 = (beta[1], beta[2]))
end # module LocalDG

# ------------------------------------------------------------------------- C structs
struct LdgParams           # stefan_ldg_params
    nx::Int32; ny::Int32
    hx::Float64; hy::Float64
    order::Int32; numqp::Int32
    k1::Float64; k2::Float64
    interiorpenalty::Float64; negboundarypenalty::Float64; posboundarypenalty::Float64
    V0x::Float64; V0y::Float64
    interfacepenalty::Float64; aligned_interface::Int32
end

mutable struct LDGOperator
    handle::Ptr{Cvoid}
    ndofs::Int
    function LDGOperator(mesh::UniformMesh, sp::Dict{Symbol,Any})
        h = element_size(mesh)
        p = Ref(LdgParams(mesh.nelements[1], mesh.nelements[2], h[1], h[2], sp[:order], sp[:numqp], sp[:k1], sp[:k2],
                          sp[:interiorpenalty], sp[:negboundarypenalty], sp[:posboundarypenalty], sp[:V0]...,
                          get(sp, :interfacepenalty, 0.0), haskey(sp, :interfacepenalty) ? 1 : 0))
        out = Ref{Ptr{Cvoid}}(C_NULL)
        check(ccall((:stefan_ldg_create, lib), Cint,
                    (Ref{LdgParams}, Ptr{Int8}, Ptr{Int8}, Ptr{Int8}, Ref{Ptr{Cvoid}}),
                    p, mesh.cellsign, C_NULL, C_NULL, out))
        op = new(out[], ccall((:stefan_ldg_ndofs, lib), Int64, (Ptr{Cvoid},), out[]))
        finalizer(o -> ccall((:stefan_ldg_destroy, lib), Cint, (Ptr{Cvoid},), o.handle), op)
        op
    end
end

"y = A x on the device (replaces `sparse_operator(sysmatrix, mesh, 3) * x`)."
function Base.:*(A::LDGOperator, x::CuVector{Float64})
    y = similar(x)
    check(ccall((:stefan_ldg_apply, lib), Cint,
                (Ptr{Cvoid}, CuPtr{Float64}, CuPtr{Float64}, CuPtr{Float64}, CuPtr{Float64}, Ptr{Cvoid}),
                A.handle, x, y, CU_NULL, CU_NULL, CUDA.stream().handle))
    y
end

"`matrix \\ rhs` for the local-DG system on the device: Schur complement on T + BiCGStab (stefan_ldg_solve)."
function Base.:\(A::LDGOperator, b::CuVector{Float64}; tol = 1e-12, maxiter = 20000)
    x = similar(b)
    res = Ref(CgResult(0, 0.0, 0.0, 0))
    check(ccall((:stefan_ldg_solve, lib), Cint,
                (Ptr{Cvoid}, CuPtr{Float64}, CuPtr{Float64}, Ref{CgParams}, Ref{CgResult}, Ptr{Cvoid}),
                A.handle, b, x, Ref(CgParams(maxiter, tol, 0, 10)), res, CUDA.stream().handle))
    res[].converged == 1 || @warn "BiCGStab did not converge" res[]
    x
end

# ------------------------------------------------------------------ element-block path
# Arbitrary per-cell / per-face quadratures (cut and merged cells): flat device arrays in the
# layout of `stefan_blocks_data` (include/stefan_b200.h); see INTEGRATION.md, "Cut-cell data".
struct BlocksDims          # stefan_blocks_dims
    ncells::Int32; order::Int32; nqv::Int32; nslots::Int32; nqf::Int32
end
struct BlocksData          # stefan_blocks_data (device pointers)
    vol_pts::CuPtr{Float64}; vol_wts::CuPtr{Float64}; cell_jac::CuPtr{Float64}; cell_k::CuPtr{Float64}
    slot_nbr::CuPtr{Int32}
    face_pts::CuPtr{Float64}; face_nbr_pts::CuPtr{Float64}; face_wts::CuPtr{Float64}; face_normals::CuPtr{Float64}
    slot_pen::CuPtr{Float64}; slot_lambda::CuPtr{Float64}
end
mutable struct BlockSystem
    handle::Ptr{Cvoid}
    ndofs::Int
    function BlockSystem(dims::BlocksDims)
        out = Ref{Ptr{Cvoid}}(C_NULL)
        check(ccall((:stefan_blocks_create, lib), Cint, (Ref{BlocksDims}, Ref{Ptr{Cvoid}}), Ref(dims), out))
        b = new(out[], ccall((:stefan_blocks_ndofs, lib), Int64, (Ptr{Cvoid},), out[]))
        finalizer(o -> ccall((:stefan_blocks_destroy, lib), Cint, (Ptr{Cvoid},), o.handle), b)
        b
    end
end
"Boundary form of the block path: 0 = current, 1 = legacy_boundary (what the reference's stored cut-cell IP files were
produced with; stefan_blocks_set_variant).  Call before `assemble!`."
set_variant!(b::BlockSystem, variant::Integer) =
    check(ccall((:stefan_blocks_set_variant, lib), Cint, (Ptr{Cvoid}, Int32), b.handle, variant))
assemble!(b::BlockSystem, data::BlocksData) =
    check(ccall((:stefan_blocks_assemble, lib), Cint, (Ptr{Cvoid}, Ref{BlocksData}, Ptr{Cvoid}),
                b.handle, Ref(data), CUDA.stream().handle))
function Base.:*(b::BlockSystem, x::CuVector{Float64})
    y = similar(x)
    check(ccall((:stefan_blocks_apply, lib), Cint, (Ptr{Cvoid}, CuPtr{Float64}, CuPtr{Float64}, Ptr{Cvoid}),
                b.handle, x, y, CUDA.stream().handle))
    y
end
"rhs of assemble_interior_penalty_rhs! + assemble_two_phase_source! + assemble_robin_source! for block data
(f_qp [ncells][nqv] source samples, g_qp [ncells][nslots][nqf] boundary data; stefan_blocks_rhs)."
function rhs_vector(b::BlockSystem, data::BlocksData, f_qp::CuVector{Float64}, g_qp::CuVector{Float64}, Tm)
    out = CUDA.zeros(Float64, b.ndofs)
    check(ccall((:stefan_blocks_rhs, lib), Cint,
                (Ptr{Cvoid}, Ref{BlocksData}, CuPtr{Float64}, CuPtr{Float64}, Float64, CuPtr{Float64}, Ptr{Cvoid}),
                b.handle, Ref(data), f_qp, g_qp, Tm, out, CUDA.stream().handle))
    out
end
"`matrix \\ rhs` for the (symmetric) cut-cell IP system: block-Jacobi CG on the device (stefan_blocks_cg_solve)."
function Base.:\(b::BlockSystem, rhs::CuVector{Float64}; tol = 1e-12, maxiter = 100000)
    x = CUDA.zeros(Float64, length(rhs))
    res = Ref(CgResult(0, 0.0, 0.0, 0))
    check(ccall((:stefan_blocks_cg_solve, lib), Cint,
                (Ptr{Cvoid}, CuPtr{Float64}, CuPtr{Float64}, Ref{CgParams}, Ref{CgResult}, Ptr{Cvoid}),
                b.handle, rhs, x, Ref(CgParams(maxiter, tol, 1, 10)), res, CUDA.stream().handle))
    res[].converged == 1 || @warn "CG did not converge" res[]
    x
end
"mesh_L2_error / integral_norm_on_mesh over the solution cells' volume points (stefan_blocks_l2_error)."
function l2_error(b::BlockSystem, data::BlocksData, u::CuVector{Float64}, uex_qp::CuVector{Float64})
    out = CUDA.zeros(Float64, 2)
    check(ccall((:stefan_blocks_l2_error, lib), Cint,
                (Ptr{Cvoid}, Ref{BlocksData}, CuPtr{Float64}, CuPtr{Float64}, CuPtr{Float64}, Ptr{Cvoid}),
                b.handle, Ref(data), u, uex_qp, out, CUDA.stream().handle))
    Tuple(Array(out))
end

# ------------------------------------------------- local DG on element blocks (cut meshes; stefan_ldgblocks_*)
"assemble_LDG_linear_system! (LDG_assembly.jl:1-91) for per-cell / per-face quadrature data: 3 dofs per node
(T, q_x, q_y); `data.slot_pen` holds every slot's penalty (interior / interface / boundary by sign(V0 . n))."
mutable struct LDGBlockSystem
    handle::Ptr{Cvoid}
    ndofs::Int
    function LDGBlockSystem(dims::BlocksDims)
        out = Ref{Ptr{Cvoid}}(C_NULL)
        check(ccall((:stefan_ldgblocks_create, lib), Cint, (Ref{BlocksDims}, Ref{Ptr{Cvoid}}), Ref(dims), out))
        b = new(out[], ccall((:stefan_ldgblocks_ndofs, lib), Int64, (Ptr{Cvoid},), out[]))
        finalizer(o -> ccall((:stefan_ldgblocks_destroy, lib), Cint, (Ptr{Cvoid},), o.handle), b)
        b
    end
end
assemble!(b::LDGBlockSystem, data::BlocksData, V0) =
    check(ccall((:stefan_ldgblocks_assemble, lib), Cint, (Ptr{Cvoid}, Ref{BlocksData}, Float64, Float64, Ptr{Cvoid}),
                b.handle, Ref(data), V0[1], V0[2], CUDA.stream().handle))
function Base.:*(b::LDGBlockSystem, x::CuVector{Float64})
    y = similar(x)
    check(ccall((:stefan_ldgblocks_apply, lib), Cint, (Ptr{Cvoid}, CuPtr{Float64}, CuPtr{Float64}, Ptr{Cvoid}),
                b.handle, x, y, CUDA.stream().handle))
    y
end
"assemble_LDG_rhs! (LDG_assembly.jl:93-131) + the two-phase source for block data (stefan_ldgblocks_rhs)."
function rhs_vector(b::LDGBlockSystem, data::BlocksData, f_qp::CuVector{Float64}, g_qp::CuVector{Float64})
    out = CUDA.zeros(Float64, b.ndofs)
    check(ccall((:stefan_ldgblocks_rhs, lib), Cint,
                (Ptr{Cvoid}, Ref{BlocksData}, CuPtr{Float64}, CuPtr{Float64}, CuPtr{Float64}, Ptr{Cvoid}),
                b.handle, Ref(data), f_qp, g_qp, out, CUDA.stream().handle))
    out
end
"`reshape(matrix \\ rhs, 3, :)` of LDG_test_cylindrical_bc_convergence.jl:159 on the device (Schur complement + BiCGStab)."
function Base.:\(b::LDGBlockSystem, rhs::CuVector{Float64}; tol = 1e-12, maxiter = 200000)
    x = CUDA.zeros(Float64, length(rhs))
    res = Ref(CgResult(0, 0.0, 0.0, 0))
    check(ccall((:stefan_ldgblocks_solve, lib), Cint,
                (Ptr{Cvoid}, CuPtr{Float64}, CuPtr{Float64}, Ref{CgParams}, Ref{CgResult}, Ptr{Cvoid}),
                b.handle, rhs, x, Ref(CgParams(maxiter, tol, 0, 10)), res, CUDA.stream().handle))
    res[].converged == 1 || @warn "BiCGStab did not converge" res[]
    reshape(x, 3, :)
end

# ------------------------------------------------- point evaluation / front velocity (IP_postprocess.jl)
"InteriorPenalty.interpolate_at_reference_points (IP_postprocess.jl:30-52): refpoints 2 x numpts, refcellids 1-based
solution-cell ids (the caller picks the cell of the wanted level-set sign)."
function interpolate_at_reference_points(nodalvalues::CuVector{Float64}, basis, refpoints::Matrix{Float64}, refcellids)
    n = size(refpoints, 2)
    pts, ids, out = CuArray(vec(refpoints)), CuArray(Int32.(refcellids .- 1)), CUDA.zeros(Float64, n)
    check(ccall((:stefan_interpolate_at_reference_points, lib), Cint,
                (Int32, CuPtr{Float64}, CuPtr{Float64}, CuPtr{Int32}, Int64, CuPtr{Float64}, Ptr{Cvoid}),
                basis.order, nodalvalues, pts, ids, n, out, CUDA.stream().handle))
    out
end
"InteriorPenalty.gradient_at_reference_points (IP_postprocess.jl:1-28): 2 x numpts physical gradients."
function gradient_at_reference_points(nodalvalues::CuVector{Float64}, basis, refpoints::Matrix{Float64}, refcellids, jacobian)
    n = size(refpoints, 2)
    pts, ids, out = CuArray(vec(refpoints)), CuArray(Int32.(refcellids .- 1)), CUDA.zeros(Float64, 2, n)
    check(ccall((:stefan_gradient_at_reference_points, lib), Cint,
                (Int32, CuPtr{Float64}, CuPtr{Float64}, CuPtr{Int32}, Int64, Float64, Float64, CuPtr{Float64}, Ptr{Cvoid}),
                basis.order, nodalvalues, pts, ids, n, jacobian[1], jacobian[2], out, CUDA.stream().handle))
    out
end

# ------------------------------------------------------------------------- C structs (IP)
struct IpdgParams          # stefan_ipdg_params
    nx::Int32; ny::Int32
    hx::Float64; hy::Float64
    order::Int32; numqp::Int32
    k1::Float64; k2::Float64
    penalty::Float64
    lambda::Float64; Tm::Float64
    variant::Int32; terms::Int32
end

mutable struct IPOperator
    handle::Ptr{Cvoid}
    ndofs::Int
    function IPOperator(mesh::UniformMesh, order, numqp, k1, k2, penalty, lambda, Tm, variant, terms)
        h = element_size(mesh)
        p = Ref(IpdgParams(mesh.nelements[1], mesh.nelements[2], h[1], h[2], order, numqp, k1, k2,
                           penalty, lambda, Tm, variant, terms))
        out = Ref{Ptr{Cvoid}}(C_NULL)
        check(ccall((:stefan_ipdg_create, lib), Cint,
                    (Ref{IpdgParams}, Ptr{Int8}, Ptr{Int8}, Ptr{Int8}, Ref{Ptr{Cvoid}}),
                    p, mesh.cellsign, C_NULL, C_NULL, out))
        op = new(out[], ccall((:stefan_ipdg_ndofs, lib), Int64, (Ptr{Cvoid},), out[]))
        finalizer(o -> ccall((:stefan_ipdg_destroy, lib), Cint, (Ptr{Cvoid},), o.handle), op)
        op
    end
end

"y = A x on the device (replaces `sparse_operator(...) * x`)."
function Base.:*(A::IPOperator, x::CuVector{Float64})
    y = similar(x)
    check(ccall((:stefan_ipdg_apply, lib), Cint,
                (Ptr{Cvoid}, CuPtr{Float64}, CuPtr{Float64}, CuPtr{Float64}, CuPtr{Float64}, Int32, Ptr{Cvoid}),
                A.handle, x, y, CU_NULL, CU_NULL, 0, CUDA.stream().handle))
    y
end

"unew = u + dt Minv (b - A u): the fused explicit step, one pass of the apply kernel (stefan_ipdg_step)."
function step!(unew::CuVector{Float64}, A::IPOperator, u::CuVector{Float64}, b::CuVector{Float64}, dt)
    check(ccall((:stefan_ipdg_step, lib), Cint,
                (Ptr{Cvoid}, CuPtr{Float64}, CuPtr{Float64}, CuPtr{Float64}, Float64, CuPtr{Float64}, CuPtr{Float64},
                 Ptr{Cvoid}, Ptr{Cvoid}),
                A.handle, u, unew, b, dt, CU_NULL, CU_NULL, C_NULL, CUDA.stream().handle))
    unew
end

"y = A x for host vectors (pinned memory recommended); copies are pipelined inside the call."
function apply_host!(y::Vector{Float64}, A::IPOperator, x::Vector{Float64})
    check(ccall((:stefan_ipdg_apply_host, lib), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Int32),
                A.handle, x, y, 0))
    y
end

struct CgParams            # stefan_cg_params
    max_iterations::Int32
    rel_tolerance::Float64
    preconditioner::Int32
    check_every::Int32
end
struct CgResult            # stefan_cg_result
    iterations::Int32
    residual_norm::Float64
    rhs_norm::Float64
    converged::Int32
end

"`matrix \\ rhs` on the device: block-Jacobi preconditioned CG over the matrix-free apply."
function Base.:\(A::IPOperator, b::CuVector{Float64}; tol = 1e-12, maxiter = 10000)
    x = CUDA.zeros(Float64, length(b))
    res = Ref(CgResult(0, 0.0, 0.0, 0))
    check(ccall((:stefan_ipdg_cg_solve, lib), Cint,
                (Ptr{Cvoid}, CuPtr{Float64}, CuPtr{Float64}, Ref{CgParams}, Ref{CgResult}, Ptr{Cvoid}),
                A.handle, b, x, Ref(CgParams(maxiter, tol, 1, 10)), res, CUDA.stream().handle))
    res[].converged == 1 || @warn "CG did not converge" res[]
    x
end

"""
The same solve across column slabs, one process per GPU (stefan_ipdg_cg_solve_slab).  `exchange(v, stream)` must fill
`halo_lo` / `halo_hi` with the neighbours' edge columns of the slab vector at device pointer `v` (e.g. CUDA-aware
`MPI.Sendrecv!`), `allreduce(dev, count, stream)` must sum `count` device doubles over the ranks in place
(`MPI.Allreduce!`); both are ordered on the raw stream they receive and return nothing.
"""
function solve_slab(A::IPOperator, b::CuVector{Float64}, halo_lo, halo_hi, exchange, allreduce;
                    tol = 1e-12, maxiter = 10000)
    x = CUDA.zeros(Float64, length(b))
    res = Ref(CgResult(0, 0.0, 0.0, 0))
    ex(ctx, v, st) = (exchange(v, st); Cint(0))
    ar(ctx, d, n, st) = (allreduce(d, n, st); Cint(0))
    cex = @cfunction($ex, Cint, (Ptr{Cvoid}, CuPtr{Float64}, Ptr{Cvoid}))
    car = @cfunction($ar, Cint, (Ptr{Cvoid}, CuPtr{Float64}, Int32, Ptr{Cvoid}))
    lo = halo_lo === nothing ? CU_NULL : pointer(halo_lo)
    hi = halo_hi === nothing ? CU_NULL : pointer(halo_hi)
    GC.@preserve cex car check(ccall((:stefan_ipdg_cg_solve_slab, lib), Cint,
                (Ptr{Cvoid}, CuPtr{Float64}, CuPtr{Float64}, Ref{CgParams}, Ref{CgResult}, CuPtr{Float64}, CuPtr{Float64},
                 Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}),
                A.handle, b, x, Ref(CgParams(maxiter, tol, 1, 10)), res, lo, hi, cex, car, C_NULL, CUDA.stream().handle))
    res[].converged == 1 || @warn "CG did not converge" res[]
    x
end

"(||u_h - u_ex||, ||u_ex||) in L2 with the cell quadrature (`mesh_L2_error`, useful_routines.jl:95-133);
`uex_qp` = exact solution at the cell quadrature points, [ncells][numqp^2]."
function l2_error(A::IPOperator, u::CuVector{Float64}, uex_qp::CuVector{Float64})
    out = CUDA.zeros(Float64, 2)
    check(ccall((:stefan_ipdg_l2_error, lib), Cint,
                (Ptr{Cvoid}, CuPtr{Float64}, CuPtr{Float64}, CuPtr{Float64}, Ptr{Cvoid}),
                A.handle, u, uex_qp, out, CUDA.stream().handle))
    Tuple(Array(out))
end

"The COO seam: (rows, cols, vals) with Julia (1-based) indices, ready for `sparse(rows, cols, vals, N, N)`."
function coo(A::IPOperator)
    n = ccall((:stefan_ipdg_coo_size, lib), Int64, (Ptr{Cvoid},), A.handle)
    rows, cols, vals = CUDA.zeros(Int64, n), CUDA.zeros(Int64, n), CUDA.zeros(Float64, n)
    nnz = Ref{Int64}(0)
    check(ccall((:stefan_ipdg_assemble_coo, lib), Cint,
                (Ptr{Cvoid}, CuPtr{Int64}, CuPtr{Int64}, CuPtr{Float64}, Int64, Int32, Ref{Int64}, Ptr{Cvoid}),
                A.handle, rows, cols, vals, n, 1, nnz, CUDA.stream().handle))
    Array(rows), Array(cols), Array(vals)
end

"CutCellDG.sparse_operator(sysmatrix, mesh, dofspernode) -> device operator."
function sparse_operator(s::SystemMatrix, mesh::UniformMesh, dofspernode::Integer)
    if s.kind == :ip
        dofspernode == 1 || error("interior penalty has 1 dof per node")
        sp = s.spec
        return IPOperator(mesh, sp[:order], sp[:numqp], sp[:k1], sp[:k2], get(sp, :penalty, 0.0),
                          get(sp, :lambda, 0.0), 0.0, 0, s.terms)
    end
    if s.kind == :ldg
        dofspernode == 3 || error("local DG has 3 dofs per node")
        return LDGOperator(mesh, s.spec)
    end
    # :dg1d -> stefan_dg1d_create / stefan_dg1d_apply, same pattern
    # (see INTEGRATION.md and the Python twin stefanproblem_b200/dg1d.py)
    error("nothing assembled (kind = $(s.kind))")
end

# ------------------------------------------------------------------- FiniteDifference
module FiniteDifference
import ..lib, ..check
using CUDA
mutable struct SystemMatrix
    phaseid::Vector{Int8}
    stepsize::Float64
    dirichlet::Bool
    continuity::Vector{Tuple{Float64,Float64,Int}}
    SystemMatrix() = new(Int8[], 0.0, false, Tuple{Float64,Float64,Int}[])
end
phase_id(phi) = Int8[p < 0 ? 1 : 2 for p in phi]                      # finite-difference.jl:92-103
function assemble_laplace_operator!(m, phaseid, stepsize)             # :71-85
    m.phaseid = Int8.(phaseid); m.stepsize = stepsize
end
assemble_dirichlet_boundary_condition!(m, numnodes) = (m.dirichlet = true)   # :87-90
assemble_interface_continuity!(m, r, s, nodeid) = push!(m.continuity, (r, s, nodeid))  # :111-116
mutable struct FDOperator
    handle::Ptr{Cvoid}
    n::Int
end
function sparse_operator(m::SystemMatrix, ndofs)                      # :105-109
    out = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:stefan_fd_create, lib), Cint,
                (Int64, Float64, Ptr{Int8}, Int32, Ptr{Int8}, Ptr{Int8}, Ref{Ptr{Cvoid}}),
                ndofs, m.stepsize, m.phaseid, m.dirichlet, C_NULL, C_NULL, out))
    for (r, s, nodeid) in m.continuity
        check(ccall((:stefan_fd_add_interface_continuity, lib), Cint, (Ptr{Cvoid}, Float64, Float64, Int64),
                    out[], r, s, nodeid - 1))
    end
    op = FDOperator(out[], ndofs)
    finalizer(o -> ccall((:stefan_fd_destroy, lib), Cint, (Ptr{Cvoid},), o.handle), op)
    op
end
function Base.:*(K::FDOperator, u::CuVector{Float64})
    y = similar(u)
    check(ccall((:stefan_fd_apply, lib), Cint,
                (Ptr{Cvoid}, CuPtr{Float64}, CuPtr{Float64}, CuPtr{Float64}, CuPtr{Float64}, Ptr{Cvoid}),
                K.handle, u, y, CU_NULL, CU_NULL, CUDA.stream().handle))
    y
end
end # module FiniteDifference

end # module StefanB200
