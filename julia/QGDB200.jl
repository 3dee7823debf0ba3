# QGDB200.jl -- Julia-side binding of libqgd_b200.so (include/qgd_b200.h) for QuantumGateDesign.jl.
#
# NOT EXECUTED in this repository's CI: neither the build image nor the GPU box has Julia
# (SURVEY section 0.10).  The same C ABI is exercised on every run through Python ctypes
# (quantumgatedesign.jl_b200/_abi.py, tests/).  This file is what a maintainer of the reference
# would `include` from src/QuantumGateDesign.jl after the existing includes (it needs
# SchrodingerProb, the control types, the preconditioner types and complex_to_real).
#
# The style follows the reference's own ccall of its Fortran library
# (src/Controls/FortranBSpline.jl:1, 257-265).

using SparseArrays

const libqgd = joinpath(dirname(pathof(@__MODULE__)), "..", "deps", "libqgd_b200.so")

# ---- struct mirrors (field order and types == include/qgd_b200.h) -------------------------------------
struct QgdMatrix
    kind::Int32
    reserved::Int32
    nrows::Int64
    ncols::Int64
    dense::Ptr{Float64}
    nnz::Int64
    colptr::Ptr{Int64}
    rowval::Ptr{Int64}
    nzval::Ptr{Float64}
end
QgdMatrix(A::Matrix{Float64}) =
    QgdMatrix(0, 0, size(A, 1), size(A, 2), pointer(A), 0, C_NULL, C_NULL, C_NULL)
QgdMatrix(A::SparseMatrixCSC{Float64,Int64}) =
    QgdMatrix(1, 0, size(A, 1), size(A, 2), C_NULL, nnz(A), pointer(A.colptr), pointer(A.rowval), pointer(A.nzval))

struct QgdControl
    type::Int32
    reserved::Int32
    tf::Float64
    n_amplitudes::Int64
    D1::Int64
    degree::Int64
    n_basis::Int64
    n_carriers::Int64
    carrier_freqs::Ptr{Float64}
end
qgd_control(c::GRAPEControl) = QgdControl(1, 0, c.tf, c.N_amplitudes, 0, 0, 0, 0, C_NULL)
qgd_control(c::BSpline2Control) = QgdControl(2, 0, c.tf, 0, c.D1, 0, 0, 0, C_NULL)
qgd_control(c::FortranBSplineControl) = QgdControl(3, 0, c.tf, 0, 0, c.degree, c.N_basis_functions, 0, C_NULL)
function qgd_control(c::CarrierControl)
    b = qgd_control(c.base_control)
    return QgdControl(b.type, 0, b.tf, b.n_amplitudes, b.D1, b.degree, b.n_basis,
                      length(c.carrier_frequencies), pointer(c.carrier_frequencies))
end
# every other AbstractControl (GeneralBSplineControl, HermiteControl, HermiteCarrierControl, ...): no device kernel,
# the host evaluates the control protocol and passes tables (QGD_CONTROL_HOST_TABLE = 4, n_amplitudes = N_coeff)
qgd_control(c::AbstractControl) = QgdControl(4, 0, c.tf, c.N_coeff, 0, 0, 0, 0, C_NULL)
is_host_control(c::AbstractControl) = qgd_control(c).type == 4

struct QgdProblem
    N_tot_levels::Int64
    N_ess_levels::Int64
    N_initial_conditions::Int64
    N_operators::Int64
    system_sym::QgdMatrix
    system_asym::QgdMatrix
    sym_operators::Ptr{QgdMatrix}
    asym_operators::Ptr{QgdMatrix}
    u0::Ptr{Float64}
    v0::Ptr{Float64}
    guard_subspace_projector::QgdMatrix
    tf::Float64
    nsteps::Int64
    gmres_abstol::Float64
    gmres_reltol::Float64
    preconditioner::Int32
    reserved::Int32
    controls::Ptr{QgdControl}
end

struct QgdStats
    kernel_launches::Int64
    h2d_bytes::Int64
    d2h_bytes::Int64
    last_forward_ms::Float64
    last_backward_ms::Float64
    last_total_ms::Float64
    fast_path_launches::Int64
end

qgd_precond(::Type{IdentityPreconditioner}) = Int32(0)
qgd_precond(::Type{LUPreconditioner}) = Int32(1)
qgd_precond(::Type{DiagonalHamiltonianPreconditioner}) = Int32(2)

function qgd_check(rc::Integer)
    rc == 0 && return nothing
    msg = unsafe_string(ccall((:qgd_last_error, libqgd), Cstring, ()))
    rc == -1 && throw(ArgumentError(msg))
    error("qgd_b200 (status $rc): $msg")
end

# ---- handle ---------------------------------------------------------------------------------------------
mutable struct B200Handle
    ptr::Ptr{Cvoid}
    N_coeff::Int64
    function B200Handle(prob::SchrodingerProb{M,VM,P}, controls; device::Integer=-1) where {M,VM,P}
        ctrl_list = controls isa AbstractControl ? [controls] : collect(controls)
        length(ctrl_list) == prob.N_operators || throw(ArgumentError("need one control per control operator"))
        u0 = Matrix{Float64}(prob.u0); v0 = Matrix{Float64}(prob.v0)
        sym = [QgdMatrix(A) for A in prob.sym_operators]
        asym = [QgdMatrix(A) for A in prob.asym_operators]
        cds = [qgd_control(c) for c in ctrl_list]
        out = Ref{Ptr{Cvoid}}(C_NULL)
        # every array whose pointer is inside the structs must stay rooted during the call; qgd_create copies
        GC.@preserve prob u0 v0 sym asym cds ctrl_list begin
            p = QgdProblem(prob.N_tot_levels, prob.N_ess_levels, prob.N_initial_conditions, prob.N_operators,
                           QgdMatrix(prob.system_sym), QgdMatrix(prob.system_asym), pointer(sym), pointer(asym),
                           pointer(u0), pointer(v0), QgdMatrix(prob.guard_subspace_projector),
                           prob.tf, prob.nsteps, prob.gmres_abstol, prob.gmres_reltol, qgd_precond(P), 0, pointer(cds))
            qgd_check(ccall((:qgd_create, libqgd), Cint, (Ref{QgdProblem}, Cint, Ref{Ptr{Cvoid}}), p, device, out))
        end
        h = new(out[], sum(c.N_coeff for c in ctrl_list))
        finalizer(x -> ccall((:qgd_destroy, libqgd), Cint, (Ptr{Cvoid},), x.ptr), h)
        return h
    end
end

"Forward the reference's mutable knobs (examples/cnot3_optimize_gate.jl:51-55) before each call."
function sync_knobs!(h::B200Handle, prob::SchrodingerProb)
    qgd_check(ccall((:qgd_set_nsteps, libqgd), Cint, (Ptr{Cvoid}, Int64), h.ptr, prob.nsteps))
    qgd_check(ccall((:qgd_set_gmres_tolerances, libqgd), Cint, (Ptr{Cvoid}, Float64, Float64),
                    h.ptr, prob.gmres_abstol, prob.gmres_reltol))
end

const _handles = IdDict{Any,B200Handle}()
function b200_handle(prob::SchrodingerProb, controls; device::Integer=-1)
    h = get!(() -> B200Handle(prob, controls; device=device), _handles, prob)
    sync_knobs!(h, prob)
    return h
end

_ptr_or_null(::Nothing) = Ptr{Float64}(C_NULL)
_ptr_or_null(::Missing) = Ptr{Float64}(C_NULL)
_ptr_or_null(A::Array{Float64}) = pointer(A)

# ---- eval_forward!  (src/forward_evolution.jl:33-70) ---------------------------------------------------
function eval_forward_b200!(uv_history::Array{Float64,4}, prob::SchrodingerProb, controls,
        pcof::Vector{Float64}; order::Int=2, saveEveryNsteps::Int=1, forcing=missing, device::Integer=-1)
    m = div(order, 2)
    @assert size(uv_history) == (prob.real_system_size, 1 + m, 1 + div(prob.nsteps, saveEveryNsteps), prob.N_initial_conditions)
    h = b200_handle(prob, controls; device=device)
    GC.@preserve uv_history pcof forcing begin
        if any(is_host_control, controls)
            cvals, _ = control_tables(controls, pcof, prob.tf, prob.nsteps, m)
            qgd_check(ccall((:qgd_eval_forward_tables, libqgd), Cint,
                (Ptr{Cvoid}, Int64, Int32, Int64, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Int64}),
                h.ptr, 1, order, saveEveryNsteps, cvals, uv_history, C_NULL, C_NULL))
        elseif ismissing(forcing)
            qgd_check(ccall((:qgd_eval_forward, libqgd), Cint,
                (Ptr{Cvoid}, Ptr{Float64}, Int64, Int32, Int64, Ptr{Float64}, Ptr{Float64}, Ptr{Int64}),
                h.ptr, pcof, 1, order, saveEveryNsteps, uv_history, C_NULL, C_NULL))
        else  # forcing::Array{Float64,4} [2N, m, 1+nsteps, nic]  (src/forward_evolution.jl:118-129)
            qgd_check(ccall((:qgd_eval_forward_forced, libqgd), Cint,
                (Ptr{Cvoid}, Ptr{Float64}, Int64, Int32, Int64, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Int64}),
                h.ptr, pcof, 1, order, saveEveryNsteps, forcing, uv_history, C_NULL, C_NULL))
        end
    end
    return nothing
end

# ---- eval_grad_forced  (src/eval_grad_forced.jl:18-195): P forced solves batched on the device -----------
function eval_grad_forced_b200(prob::SchrodingerProb, controls, pcof::Vector{Float64}, target::AbstractMatrix{<:Number};
        order::Int=2, cost_type=:Infidelity, device::Integer=-1)
    cost_type == :Infidelity || throw("Invalid cost type: $cost_type")
    R = Matrix{Float64}(complex_to_real(target))
    h = b200_handle(prob, controls; device=device)
    grad = zeros(length(pcof))
    qgd_check(ccall((:qgd_eval_grad_forced, libqgd), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Int32, Ptr{Float64}),
        h.ptr, pcof, R, order, grad))
    return grad
end

# ---- host-evaluated controls: tables from the reference's own control protocol (Control.jl:99-149) ------
function control_tables(controls, pcof, tf, nsteps, m)
    Nc, P = length(controls), get_number_of_control_parameters(controls)
    cvals = zeros(Nc, 1 + m, 2, 1 + nsteps, 1)
    table = zeros(P, 1 + m, 2, 1 + nsteps)
    pmat = zeros(1 + m, Nc); qmat = zeros(1 + m, Nc)
    for n in 0:nsteps
        t = n * tf / nsteps
        fill_p_mat!(pmat, controls, t, pcof); fill_q_mat!(qmat, controls, t, pcof)   # Taylor-scaled p^(j)/j!
        cvals[:, :, 1, 1+n, 1] .= pmat'; cvals[:, :, 2, 1+n, 1] .= qmat'
        offset = 0
        for c in controls
            sl = offset+1:offset+c.N_coeff
            for j in 0:m
                eval_grad_p_derivative!(view(table, sl, 1+j, 1, 1+n), c, t, pcof[sl], j)
                eval_grad_q_derivative!(view(table, sl, 1+j, 2, 1+n), c, t, pcof[sl], j)
                table[sl, 1+j, :, 1+n] ./= factorial(j)
            end
            offset += c.N_coeff
        end
    end
    return cvals, table
end

function discrete_adjoint_b200_tables(prob::SchrodingerProb, controls, pcof::Vector{Float64},
        target::AbstractMatrix{<:Number}; order::Int=2, device::Integer=-1)
    R = Matrix{Float64}(complex_to_real(target))
    h = b200_handle(prob, controls; device=device)
    cvals, table = control_tables(controls, pcof, prob.tf, prob.nsteps, div(order, 2))
    grad = zeros(length(pcof)); infid = Ref(0.0); guard = Ref(0.0)
    qgd_check(ccall((:qgd_discrete_adjoint_tables, libqgd), Cint,
        (Ptr{Cvoid}, Int64, Int32, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ref{Float64}, Ref{Float64}),
        h.ptr, 1, order, cvals, table, R, grad, infid, guard))
    return grad, infid[], guard[]
end

# ---- discrete_adjoint!  (src/eval_grad_discrete_adjoint.jl:107-160) ------------------------------------
function discrete_adjoint_b200!(grad::Vector{Float64}, history, lambda_history, adjoint_forcing,
        prob::SchrodingerProb, controls, pcof::Vector{Float64}, target::AbstractMatrix{<:Number};
        order::Int=2, cost_type=:Infidelity, history_precomputed::Bool=false, device::Integer=-1)
    cost_type == :Infidelity || throw("Invalid cost type: $cost_type")
    R = Matrix{Float64}(complex_to_real(target))        # reference :126
    h = b200_handle(prob, controls; device=device)
    infid = Ref(0.0); guard = Ref(0.0)
    GC.@preserve grad history lambda_history adjoint_forcing pcof R begin
        qgd_check(ccall((:qgd_discrete_adjoint, libqgd), Cint,
            (Ptr{Cvoid}, Ptr{Float64}, Int64, Ptr{Float64}, Int32, Int32,
             Ptr{Float64}, Ref{Float64}, Ref{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64},
             Ptr{Int64}, Ptr{Int64}, Ptr{Int64}),
            h.ptr, pcof, 1, R, order, history_precomputed,
            grad, infid, guard, _ptr_or_null(history), _ptr_or_null(lambda_history), _ptr_or_null(adjoint_forcing),
            C_NULL, C_NULL, C_NULL))
    end
    return grad, infid[], guard[]
end

"Batched sweep over many control vectors (examples/optimization_with_random_pcof.jl style): pcofs [P, B]."
function discrete_adjoint_b200_batch(prob::SchrodingerProb, controls, pcofs::Matrix{Float64},
        target::AbstractMatrix{<:Number}; order::Int=2, device::Integer=-1)
    R = Matrix{Float64}(complex_to_real(target))
    h = b200_handle(prob, controls; device=device)
    P, B = size(pcofs)
    grad = zeros(P, B); infid = zeros(B); guard = zeros(B)
    qgd_check(ccall((:qgd_discrete_adjoint, libqgd), Cint,
        (Ptr{Cvoid}, Ptr{Float64}, Int64, Ptr{Float64}, Int32, Int32,
         Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64},
         Ptr{Int64}, Ptr{Int64}, Ptr{Int64}),
        h.ptr, pcofs, B, R, order, false, grad, infid, guard, C_NULL, C_NULL, C_NULL, C_NULL, C_NULL, C_NULL))
    return grad, infid, guard
end
