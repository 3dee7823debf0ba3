# QGDB200.jl -- Julia-side binding of libqgd_b200.so (include/qgd_b200.h) for QuantumGateDesign.jl.
#
# NOT EXECUTED in this repository's CI: neither the build image nor the GPU box has Julia
# (SURVEY section 0.10).  The same C ABI is exercised on every run through Python ctypes
# (quantumgatedesign.jl_b200/_abi.py, tests/) and from plain C (tests/abi_c/abi_smoke.c).  This file is what a
# maintainer of the reference would `include` from src/QuantumGateDesign.jl after the existing includes (it needs
# SchrodingerProb, the control types, the preconditioner types and complex_to_real).
#
# THE SEAM IS MULTIPLE DISPATCH, no caller changes: `gpu_prob = b200(prob)` returns a SchrodingerProb that shares every
# array with `prob` and whose preconditioner type parameter is the tag `B200{P}`.  The methods below are more specific
# than the reference's on exactly that parameter, with the reference's exact signatures
#     eval_forward!(uv_history, prob, controls, pcof; order, saveEveryNsteps, forcing)      src/forward_evolution.jl:33-37
#     discrete_adjoint!(grad, history, lambda_history, adjoint_forcing, prob, controls, pcof, target;
#                       order, cost_type, history_precomputed)                              src/eval_grad_discrete_adjoint.jl:107-115
#     eval_grad_forced(prob, controls, pcof, target; order, cost_type)                      src/eval_grad_forced.jl:18-26
# so eval_forward, discrete_adjoint, infidelity, optimize_gate (src/ipopt_optimal_control.jl:257,304), get_histories,
# estimate_timesteps_per_period ... reach the GPU unchanged when handed `gpu_prob`.  Anything NOT overridden (eval_adjoint!,
# the finite-difference gradient, plotting) keeps working on the CPU, because B200{P}(prob, order, adjoint) constructs the
# reference's own preconditioner P.
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
    collectives::Int64
end

# ---- the dispatch tag -------------------------------------------------------------------------------------
"""
    B200{P}

Preconditioner-type tag: a `SchrodingerProb{M,VM,B200{P}}` runs the gradient hot path on the B200 with the reference
preconditioner `P` (Identity / LU / DiagonalHamiltonian).  As a constructor it builds `P` itself, so CPU code paths
that were not overridden still get a working preconditioner (protocol of src/preconditioners.jl:1-28).
"""
struct B200{P<:AbstractQGDPreconditioner} <: AbstractQGDPreconditioner end
B200{P}(prob::SchrodingerProb, order::Int, adjoint::Bool=false) where {P} = P(prob, order, adjoint)

"A problem that shares every array with `prob` and dispatches the hot path to the GPU."
function b200(prob::SchrodingerProb{M,VM,P}) where {M,VM,P}
    P <: B200 && return prob
    return SchrodingerProb(prob.system_sym, prob.system_asym, prob.sym_operators, prob.asym_operators, prob.u0, prob.v0,
                           prob.guard_subspace_projector, prob.tf, prob.nsteps, prob.N_ess_levels,
                           prob.gmres_abstol, prob.gmres_reltol, B200{P})
end
const B200Prob{M,VM,P} = SchrodingerProb{M,VM,B200{P}}

qgd_precond(::Type{IdentityPreconditioner}) = Int32(0)
qgd_precond(::Type{LUPreconditioner}) = Int32(1)
qgd_precond(::Type{DiagonalHamiltonianPreconditioner}) = Int32(2)
qgd_precond(::Type{B200{P}}) where {P} = qgd_precond(P)

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

"Forward the reference's mutable knobs (examples/cnot3_optimize_gate.jl:51-55) before each call.  The C setters are
no-ops when the value is unchanged, so this does not invalidate a resident history (history_precomputed)."
function sync_knobs!(h::B200Handle, prob::SchrodingerProb)
    qgd_check(ccall((:qgd_set_nsteps, libqgd), Cint, (Ptr{Cvoid}, Int64), h.ptr, prob.nsteps))
    qgd_check(ccall((:qgd_set_gmres_tolerances, libqgd), Cint, (Ptr{Cvoid}, Float64, Float64),
                    h.ptr, prob.gmres_abstol, prob.gmres_reltol))
end

"qgd_set_option (include/qgd_b200.h), e.g. `set_option!(h, 1, 1)` for strict modified Gram-Schmidt."
set_option!(h::B200Handle, key::Integer, value::Integer) =
    qgd_check(ccall((:qgd_set_option, libqgd), Cint, (Ptr{Cvoid}, Int32, Int64), h.ptr, key, value))

# What qgd_create reads from a control: type and shape parameters (not the coefficients).
control_key(c::CarrierControl) = (:carrier, control_key(c.base_control), Tuple(c.carrier_frequencies))
control_key(c::AbstractControl) = (nameof(typeof(c)), c.N_coeff, c.tf,
                                   Tuple(getfield(c, f) for f in fieldnames(typeof(c)) if getfield(c, f) isa Integer))

# One handle per (problem object, control descriptors, device).  Weak keys: the handle (and its device memory) goes
# when the problem does; a different set of controls on the same problem gets its own handle.
const _handles = WeakKeyDict{Any,Dict{Any,B200Handle}}()
function b200_handle(prob::SchrodingerProb, controls; device::Integer=-1)
    ctrl_list = controls isa AbstractControl ? [controls] : collect(controls)
    key = (Tuple(control_key(c) for c in ctrl_list), prob.tf, Int(device))
    per_prob = get!(() -> Dict{Any,B200Handle}(), _handles, prob)
    h = get!(() -> B200Handle(prob, ctrl_list; device=device), per_prob, key)
    sync_knobs!(h, prob)
    return h
end

_ptr_or_null(::Nothing) = Ptr{Float64}(C_NULL)
_ptr_or_null(::Missing) = Ptr{Float64}(C_NULL)
_ptr_or_null(A::Array{Float64}) = pointer(A)
_ptr_or_null(A::AbstractArray{Float64}) = throw(ArgumentError("the B200 path writes into contiguous Array{Float64} buffers (got $(typeof(A)))"))

# ---- eval_forward!  (src/forward_evolution.jl:33-70): method of the reference's function for tagged problems ----
function eval_forward!(uv_history::AbstractArray{Float64,4},
        prob::SchrodingerProb{M1,M2,B200{P}}, controls, pcof::AbstractVector{<:Real};
        order::Int=2, saveEveryNsteps::Int=1,
        forcing::Union{AbstractArray{Float64,4},Missing}=missing
    ) where {M1<:AbstractMatrix{Float64},M2<:AbstractMatrix{Float64},P}
    hist = uv_history isa Array{Float64,4} ? uv_history : Array{Float64,4}(uv_history)
    eval_forward_b200!(hist, prob, controls, Vector{Float64}(pcof); order=order, saveEveryNsteps=saveEveryNsteps,
                       forcing=ismissing(forcing) ? missing : Array{Float64,4}(forcing))
    hist === uv_history || (uv_history .= hist)
    return nothing
end

function eval_forward_b200!(uv_history::Array{Float64,4}, prob::SchrodingerProb, controls,
        pcof::Vector{Float64}; order::Int=2, saveEveryNsteps::Int=1, forcing=missing, device::Integer=-1)
    m = div(order, 2)
    @assert size(uv_history) == (prob.real_system_size, 1 + m, 1 + div(prob.nsteps, saveEveryNsteps), prob.N_initial_conditions)
    h = b200_handle(prob, controls; device=device)
    GC.@preserve uv_history pcof forcing begin
        ctrl_list = controls isa AbstractControl ? [controls] : controls
        if any(is_host_control, ctrl_list)
            cvals, _ = control_tables(ctrl_list, pcof, prob.tf, prob.nsteps, m)
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
function eval_grad_forced(prob::SchrodingerProb{M,VM,B200{P}}, controls, pcof::AbstractVector{Float64},
        target::AbstractVecOrMat{<:Number}; order::Integer=2, cost_type=:Infidelity, return_forcing=false
    ) where {M<:AbstractMatrix{Float64},VM<:AbstractVecOrMat{Float64},P}
    return_forcing && throw(ArgumentError("return_forcing is not available on the B200 path (the forcing of the P forced solves is formed on the fly)"))
    return eval_grad_forced_b200(prob, controls, Vector{Float64}(pcof), target; order=Int(order), cost_type=cost_type)
end

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

# ---- discrete_adjoint!  (src/eval_grad_discrete_adjoint.jl:107-160): method of the reference's function ----------
function discrete_adjoint!(
        grad::AbstractVector{Float64}, history::AbstractArray{Float64,4},
        lambda_history::AbstractArray{Float64,4}, adjoint_forcing::AbstractArray{Float64,3},
        prob::SchrodingerProb{<:AbstractMatrix{Float64},<:AbstractMatrix{Float64},B200{P}},
        controls, pcof::AbstractVector{<:Real}, target::AbstractMatrix{<:Number};
        order=2, cost_type=:Infidelity, history_precomputed=false
    ) where {P}
    g = grad isa Vector{Float64} ? grad : Vector{Float64}(grad)
    if any(is_host_control, controls isa AbstractControl ? [controls] : controls)
        gt, _, _ = discrete_adjoint_b200_tables(prob, controls, Vector{Float64}(pcof), target; order=Int(order))
        grad .= gt
        return grad
    end
    # history_precomputed: the reference reuses the ARRAY the caller passes; here the history of the previous
    # eval_forward! / discrete_adjoint! of this (prob, controls) is still on the device and is reused when it belongs to the
    # same pcof (QGD_ESTATE otherwise, re-raised below) -- the array contents are not uploaded.
    discrete_adjoint_b200!(g, history_precomputed ? nothing : history, lambda_history, adjoint_forcing, prob, controls,
                           Vector{Float64}(pcof), target; order=Int(order), cost_type=cost_type,
                           history_precomputed=Bool(history_precomputed))
    g === grad || (grad .= g)
    return grad
end

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


# ---- multi-GPU from one Julia process (qgd_init_multi_gpu): n handles + an NCCL communicator inside the library ----
mutable struct B200MultiGPU
    ptr::Ptr{Cvoid}
    n_gpus::Int
end
function B200MultiGPU(prob::SchrodingerProb{M,VM,P}, controls, n_gpus::Integer) where {M,VM,P}
    ctrl_list = controls isa AbstractControl ? [controls] : collect(controls)
    u0 = Matrix{Float64}(prob.u0); v0 = Matrix{Float64}(prob.v0)
    sym = [QgdMatrix(A) for A in prob.sym_operators]; asym = [QgdMatrix(A) for A in prob.asym_operators]
    cds = [qgd_control(c) for c in ctrl_list]
    out = Ref{Ptr{Cvoid}}(C_NULL)
    GC.@preserve prob u0 v0 sym asym cds ctrl_list begin
        p = QgdProblem(prob.N_tot_levels, prob.N_ess_levels, prob.N_initial_conditions, prob.N_operators,
                       QgdMatrix(prob.system_sym), QgdMatrix(prob.system_asym), pointer(sym), pointer(asym),
                       pointer(u0), pointer(v0), QgdMatrix(prob.guard_subspace_projector),
                       prob.tf, prob.nsteps, prob.gmres_abstol, prob.gmres_reltol, qgd_precond(P), 0, pointer(cds))
        qgd_check(ccall((:qgd_init_multi_gpu, libqgd), Cint, (Ref{QgdProblem}, Int32, Ptr{Int32}, Ref{Ptr{Cvoid}}),
                        p, n_gpus, C_NULL, out))
    end
    mg = B200MultiGPU(out[], n_gpus)
    finalizer(x -> ccall((:qgd_multi_destroy, libqgd), Cint, (Ptr{Cvoid},), x.ptr), mg)
    return mg
end

"`shard = 0`: the columns of every control vector over the GPUs (two NCCL all-reduces per evaluation, on the device);
`shard = 1`: the control vectors over the GPUs (no collective).  pcofs [P, B] -> (grad [P, B], infidelity [B], guard [B])."
function discrete_adjoint_multi(mg::B200MultiGPU, pcofs::Matrix{Float64}, target::AbstractMatrix{<:Number};
        order::Int=2, shard::Integer=0)
    R = Matrix{Float64}(complex_to_real(target))
    P, B = size(pcofs)
    grad = zeros(P, B); infid = zeros(B); guard = zeros(B)
    qgd_check(ccall((:qgd_multi_discrete_adjoint, libqgd), Cint,
        (Ptr{Cvoid}, Ptr{Float64}, Int64, Ptr{Float64}, Int32, Int32, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}),
        mg.ptr, pcofs, B, R, order, shard, grad, infid, guard))
    return grad, infid, guard
end
