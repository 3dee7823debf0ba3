# make_reference_golden.jl -- run the REAL reference (QuantumGateDesign.jl) on the inputs exported by
# tools/export_reference_inputs.py and dump what this repository's parity tests compare:
# gradient, infidelity, guard penalty, final state, first column of lambda_history at t = 0, and GMRES iteration totals.
#
#     julia --threads 8 --project=<path of QuantumGateDesign.jl> tools/make_reference_golden.jl tests/golden/ref_inputs
#
# NOT EXECUTED in this repository: neither the build image nor the GPU box has Julia (SURVEY section 0.10).  Written
# against the reference's public API only (template: test/GradientTests/compare_gradients.jl:23-66):
#   SchrodingerProb inner constructor            src/SchrodingerProb.jl:48-164
#   GRAPEControl / BSpline2Control / FortranBSplineControl / CarrierControl constructors   src/Controls/*.jl
#   discrete_adjoint!                            src/eval_grad_discrete_adjoint.jl:107-160
#   infidelity_real, guard_penalty_real          src/infidelity.jl:7-18, 56-96
# GMRES iteration counts are not returned by the reference; they are obtained through its own extension point, the
# preconditioner protocol `P(prob, order, adjoint)` (src/preconditioners.jl:1-28): CountingPreconditioner{P} wraps the
# preconditioner of the case and counts `ldiv!` calls.  IterativeSolvers applies the left preconditioner once per
# Arnoldi step (expand!) and once per initial / restart residual (init!), so per sweep
#     iterations = ldiv_calls - (number of solves + number of restarts);
# both raw numbers are written (restarts do not occur in these cases: restart = 2N exceeds every iteration count).
# Outputs: <dir>/<case>/out_*.f64 (raw little-endian Float64, column-major) + out_meta.json; convert with
# tools/ref_golden_to_npz.py.
using QuantumGateDesign
using LinearAlgebra, SparseArrays
import QuantumGateDesign: AbstractQGDPreconditioner, IdentityPreconditioner, LUPreconditioner, DiagonalHamiltonianPreconditioner

const LDIV_FWD = Threads.Atomic{Int}(0)
const LDIV_ADJ = Threads.Atomic{Int}(0)
struct CountingPreconditioner{P<:AbstractQGDPreconditioner,Q} <: AbstractQGDPreconditioner
    P::Q
    adjoint::Bool
end
function CountingPreconditioner{P}(prob, order::Int, adjoint::Bool=false) where {P}
    inner = P(prob, order, adjoint)
    return CountingPreconditioner{P,typeof(inner)}(inner, adjoint)
end
CountingPreconditioner{P,Q}(prob, order::Int, adjoint::Bool=false) where {P,Q} = CountingPreconditioner{P}(prob, order, adjoint)
count!(c::CountingPreconditioner) = Threads.atomic_add!(c.adjoint ? LDIV_ADJ : LDIV_FWD, 1)
LinearAlgebra.ldiv!(c::CountingPreconditioner, x) = (count!(c); ldiv!(c.P, x))
LinearAlgebra.ldiv!(y, c::CountingPreconditioner, x) = (count!(c); ldiv!(y, c.P, x))
Base.:\(c::CountingPreconditioner, b) = (count!(c); c.P \ b)

# --- a JSON reader just big enough for meta.json (no package dependency): numbers, strings, bools, arrays, objects ---
function parse_json(s::AbstractString)
    i = Ref(firstindex(s))
    ws() = (while i[] <= lastindex(s) && isspace(s[i[]]); i[] = nextind(s, i[]); end)
    function val()
        ws(); c = s[i[]]
        if c == '{'
            d = Dict{String,Any}(); i[] = nextind(s, i[]); ws()
            if s[i[]] == '}'; i[] = nextind(s, i[]); return d; end
            while true
                ws(); k = val(); ws(); i[] = nextind(s, i[])   # skip ':'
                d[k] = val(); ws()
                ch = s[i[]]; i[] = nextind(s, i[])
                ch == '}' && return d
            end
        elseif c == '['
            a = Any[]; i[] = nextind(s, i[]); ws()
            if s[i[]] == ']'; i[] = nextind(s, i[]); return a; end
            while true
                push!(a, val()); ws()
                ch = s[i[]]; i[] = nextind(s, i[])
                ch == ']' && return a
            end
        elseif c == '"'
            j = nextind(s, i[]); k = j
            while s[k] != '"'; k = nextind(s, k); end
            str = String(s[j:prevind(s, k)]); i[] = nextind(s, k); return str
        else
            j = i[]
            while i[] <= lastindex(s) && !(s[i[]] in (',', '}', ']')) && !isspace(s[i[]]); i[] = nextind(s, i[]); end
            tok = s[j:prevind(s, i[])]
            tok == "true" && return true
            tok == "false" && return false
            tok == "null" && return nothing
            return occursin(r"[.eE]", tok) ? parse(Float64, tok) : parse(Int, tok)
        end
    end
    return val()
end

readf64(dir, name, dims...) = (A = Array{Float64}(undef, dims...); open(io -> read!(io, A), joinpath(dir, name * ".f64")); A)
writef64(dir, name, A) = open(io -> write(io, Array{Float64}(A)), joinpath(dir, name * ".f64"), "w")

function make_control(m)
    base = if m["type"] == "GRAPEControl"
        GRAPEControl(m["N_amplitudes"], Float64(m["tf"]))
    elseif m["type"] == "BSpline2Control"
        BSpline2Control(m["D1"], Float64(m["tf"]))
    elseif m["type"] == "FortranBSplineControl"
        FortranBSplineControl(m["degree"], m["N_basis_functions"], Float64(m["tf"]))
    else
        error("unknown control type $(m["type"])")
    end
    return haskey(m, "carrier_frequencies") ? CarrierControl(base, Float64.(m["carrier_frequencies"])) : base
end

function run_case(dir)
    meta = parse_json(read(joinpath(dir, "meta.json"), String))
    N, nic, Nc, nsteps, order = meta["N_tot_levels"], meta["N_initial_conditions"], meta["N_operators"], meta["nsteps"], meta["order"]
    conv = meta["sparse"] ? sparse : identity
    Ks = conv(readf64(dir, "system_sym", N, N)); Ss = conv(readf64(dir, "system_asym", N, N))
    symops = [conv(readf64(dir, "sym_op_$k", N, N)) for k in 1:Nc]
    asymops = [conv(readf64(dir, "asym_op_$k", N, N)) for k in 1:Nc]
    u0 = readf64(dir, "u0", N, nic); v0 = readf64(dir, "v0", N, nic)
    W = conv(readf64(dir, "guard_subspace_projector", 2N, 2N))
    P = Dict("IdentityPreconditioner" => IdentityPreconditioner, "LUPreconditioner" => LUPreconditioner,
             "DiagonalHamiltonianPreconditioner" => DiagonalHamiltonianPreconditioner)[meta["preconditioner"]]
    prob = SchrodingerProb(Ks, Ss, symops, asymops, u0, v0, W, Float64(meta["tf"]), nsteps, meta["N_ess_levels"],
                           Float64(meta["gmres_abstol"]), Float64(meta["gmres_reltol"]), CountingPreconditioner{P})
    controls = [make_control(m) for m in meta["controls"]]
    pcof = vec(readf64(dir, "pcof", meta["N_coeff"]))
    target = readf64(dir, "target_re", N, nic) + im * readf64(dir, "target_im", N, nic)

    m = div(order, 2)
    grad = zeros(length(pcof))
    history = zeros(2N, 1 + m, 1 + nsteps, nic)
    lambda_history = zeros(2N, 1 + m, 1 + nsteps, nic)
    adjoint_forcing = zeros(2N, 1 + nsteps, nic)
    LDIV_FWD[] = 0; LDIV_ADJ[] = 0
    discrete_adjoint!(grad, history, lambda_history, adjoint_forcing, prob, controls, pcof, target; order=order)
    ldiv_fwd, ldiv_adj = LDIV_FWD[], LDIV_ADJ[]
    final_state = history[:, 1, end, :]
    R = QuantumGateDesign.complex_to_real(target)
    infid = infidelity_real(final_state, R, prob.N_ess_levels)
    guard = guard_penalty_real(history, prob.tf / nsteps, prob.tf, prob.guard_subspace_projector)

    writef64(dir, "out_grad", grad); writef64(dir, "out_final_state", final_state)
    writef64(dir, "out_lambda0", lambda_history[:, 1, :, :])
    writef64(dir, "out_scalars", [infid, guard, Float64(ldiv_fwd), Float64(ldiv_adj), Float64(nsteps * nic), Float64(nsteps * nic)])
    open(joinpath(dir, "out_meta.json"), "w") do io
        print(io, "{\"julia\": \"$(VERSION)\", \"threads\": $(Threads.nthreads()), \"ldiv_fwd\": $ldiv_fwd, \"ldiv_adj\": $ldiv_adj, ",
              "\"solves_fwd\": $(nsteps * nic), \"solves_adj\": $(nsteps * nic), \"infidelity\": $infid, \"guard_penalty\": $guard}")
    end
    println(meta["name"], ": |grad| = ", norm(grad), ", infidelity = ", infid, ", GMRES iterations fwd/adj = ",
            ldiv_fwd - nsteps * nic, " / ", ldiv_adj - nsteps * nic)
end

root = length(ARGS) >= 1 ? ARGS[1] : joinpath(@__DIR__, "..", "tests", "golden", "ref_inputs")
for name in sort(readdir(root))
    isfile(joinpath(root, name, "meta.json")) && run_case(joinpath(root, name))
end
