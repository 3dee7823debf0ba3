"""The named benchmark / parity configurations of BASELINE.md section 3 (synthetic, concrete).

C1  examples/cnot2_optimization.jl:10-65      CNOT2, order 4, FortranBSpline(2, 10)
C2  examples/timestep_estimation.jl:3-28 physics + declared controls/target (cnot3_setup.jl is
    missing from the reference): CNOT3, order 8, Carrier(BSpline2(10), 3 carriers) x 3
C4  dense random N (scaled so dt*||A|| ~ 1), FortranBSpline degree 8 with 2 carriers
C5  order-12 convergence sweep inputs (random N=4; src/Tests/test_convergence.jl:83-93 shape)
Julia's MersenneTwister streams are not reproducible here; numpy `default_rng(seed)` is used
wherever the reference draws random numbers (stated in BASELINE.md).
"""
from __future__ import annotations

import numpy as np

from .controls import BSpline2Control, CarrierControl, FortranBSplineControl, get_number_of_control_parameters
from .problem import (DiagonalHamiltonianPreconditioner, DispersiveProblem, IdentityPreconditioner, SchrodingerProb,
                      create_gate)

TWO_PI = 2.0 * np.pi


def cnot2(nsteps=100, tf=100.0, gmres_tol=1e-10, seed=1):
    """C1 -> (prob, controls, pcof, target, order)"""
    xa, xb, xab = 2 * 0.1099, 2 * 0.1126, 1e-2
    freqs = TWO_PI * np.array([4.10595, 4.81526])
    kerr = TWO_PI * np.array([[xa, xab], [xab, xb]])
    prob = DispersiveProblem((2, 2), (2, 2), freqs, freqs, kerr, tf, nsteps, sparse_rep=True,
                             gmres_abstol=gmres_tol, gmres_reltol=gmres_tol)
    controls = [FortranBSplineControl(2, 10, tf) for _ in range(prob.N_operators)]
    P = get_number_of_control_parameters(controls)
    pcof = 1e-2 * (0.5 - np.random.default_rng(seed).random(P))
    target = np.eye(prob.N_tot_levels, prob.N_initial_conditions)
    return prob, controls, pcof, target, 4


def cnot3_physics():
    xa, xb = 2 * 0.1099, 2 * 0.1126
    xs = 0.002494 ** 2 / xa
    xab = 1e-6
    xas, xbs = np.sqrt(xa * xs), np.sqrt(xb * xs)
    freqs = TWO_PI * np.array([4.10595, 4.81526, 7.8447])
    kerr = TWO_PI * np.array([[xa, xab, xas], [xab, xb, xbs], [xas, xbs, xs]])
    return freqs, kerr


def cnot3_pcof(P, seed=0):
    return 0.04 * (np.random.default_rng(seed).random(P) - 0.5)


def cnot3(nsteps=550, tf=550.0, gmres_tol=1e-12, seed=0, subsystem_sizes=(4, 4, 4), D1=10):
    """C2 -> (prob, controls, pcof, target, order).  Smaller `subsystem_sizes` / `nsteps`
    give the reduced parity cases."""
    freqs, kerr = cnot3_physics()
    ess = (2, 2, 2)
    prob = DispersiveProblem(subsystem_sizes, ess, freqs, freqs, kerr, tf, nsteps, sparse_rep=True,
                             gmres_abstol=gmres_tol, gmres_reltol=gmres_tol,
                             preconditioner_type=DiagonalHamiltonianPreconditioner)
    # carriers per operator k: {0, -xi_kl, -xi_kl'} (cross-Kerr detunings with the two other subsystems)
    controls = []
    for k in range(3):
        others = [l for l in range(3) if l != k]
        w = [0.0] + [-kerr[k, l] for l in others]
        controls.append(CarrierControl(BSpline2Control(D1, tf), w))
    P = get_number_of_control_parameters(controls)
    pcof = cnot3_pcof(P, seed)
    # target = I_2 (x) CNOT on the essential states: |a,1,0> <-> |a,1,1>
    pairs = []
    for a in (0, 1):
        pairs.append(((a, 1, 0), (a, 1, 1)))
        pairs.append(((a, 1, 1), (a, 1, 0)))
    target = create_gate(subsystem_sizes, ess, pairs)
    return prob, controls, pcof, target, 8


def dense_random(N=16, nic=None, Nc=2, order=10, nsteps=20, n_basis=20, degree=8, carriers=(0.0, 0.7), seed=7,
                 gmres_tol=1e-13, preconditioner_type=IdentityPreconditioner, dt_norm=1.0):
    """C4-style dense problem (random symmetric / antisymmetric operators, random_problem.jl:1-35),
    scaled so that dt*||A_drift||_2 ~ dt_norm with tf = nsteps (dt = 1)."""
    rng = np.random.default_rng(seed)
    nic = N if nic is None else nic

    def rsym():
        r = rng.random((N, N))
        return r + r.T

    def rasym():
        r = rng.random((N, N))
        return r - r.T

    Ks, Ss = rsym(), rasym()
    scale = dt_norm / np.linalg.norm(Ks + 1j * Ss, 2)
    Ks, Ss = Ks * scale, Ss * scale
    sym_ops = [rsym() * scale * 0.5 for _ in range(Nc)]
    asym_ops = [rasym() * scale * 0.5 for _ in range(Nc)]
    U0 = np.eye(N, nic) + 0j
    tf = float(nsteps)
    prob = SchrodingerProb.from_hamiltonian(Ks + 1j * Ss, sym_ops, asym_ops, U0, tf, nsteps, N,
                                            gmres_abstol=gmres_tol, gmres_reltol=gmres_tol,
                                            preconditioner_type=preconditioner_type)
    base = FortranBSplineControl(degree, n_basis, tf)
    controls = [CarrierControl(base, list(carriers)) if carriers else base for _ in range(Nc)]
    P = get_number_of_control_parameters(controls)
    pcof = rng.random(P) - 0.5
    Q, _ = np.linalg.qr(rng.standard_normal((N, N)) + 1j * rng.standard_normal((N, N)))
    target = Q[:, :nic]
    return prob, controls, pcof, target, order
