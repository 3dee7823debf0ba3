"""Host-side mirror of the reference's problem type and the constructors the benchmark needs.

`SchrodingerProb` carries the same fields as the reference's mutable struct
(src/SchrodingerProb.jl:25-41) and validates like its inner constructor (:73-154).  The
problem constructors restate src/ProblemConstructors/{multi_qudit_systems,rabi_oscillator,
random_problem}.jl; they are host-side setup, nothing here is timed.
"""
from __future__ import annotations

import itertools
from typing import Optional, Sequence

import numpy as np
import scipy.sparse as sp

from ._abi import QGD_PRECOND_DIAGONAL, QGD_PRECOND_IDENTITY, QGD_PRECOND_LU

# preconditioner "types" (src/preconditioners.jl): plain enum values on this side of the ABI
IdentityPreconditioner = QGD_PRECOND_IDENTITY
LUPreconditioner = QGD_PRECOND_LU
DiagonalHamiltonianPreconditioner = QGD_PRECOND_DIAGONAL


def _dense(M):
    return M.toarray() if sp.issparse(M) else np.asarray(M)


def _eq(A, B):
    return np.array_equal(_dense(A), _dense(B))


class SchrodingerProb:
    """Mirror of `SchrodingerProb{M,VM,P}`.

    Either call with the real-split operators (inner constructor, src/SchrodingerProb.jl:51-64)::

        SchrodingerProb(system_sym, system_asym, sym_operators, asym_operators, u0, v0,
                        guard_subspace_projector, tf, nsteps, N_ess_levels,
                        gmres_abstol, gmres_reltol, preconditioner_type)

    or use `SchrodingerProb.from_hamiltonian` for the outer constructor (:167-233).
    """

    def __init__(self, system_sym, system_asym, sym_operators, asym_operators, u0, v0,
                 guard_subspace_projector, tf, nsteps, N_ess_levels,
                 gmres_abstol=1e-10, gmres_reltol=1e-10, preconditioner_type=IdentityPreconditioner):
        N = system_sym.shape[0]
        if system_sym.shape[0] != system_sym.shape[1]:
            raise ValueError("Real part of system Hamiltonian is not square.")
        if not _eq(system_sym, system_sym.T):
            raise ValueError("Real part of system Hamiltonian is not symmetric.")
        for i, op in enumerate(sym_operators):
            if not _eq(op, op.T):
                raise ValueError(f"Symmetric operator {i + 1} is not symmetric.")
        if not _eq(system_asym, -system_asym.T):
            raise ValueError("Imaginary part of system Hamiltonian is not anti-symmetric.")
        for i, op in enumerate(asym_operators):
            if not _eq(op, -op.T):
                raise ValueError(f"Anti-symmetric operator {i + 1} is not anti-symmetric.")
        if system_asym.shape != system_sym.shape:
            raise ValueError("Size of imaginary part of Hamiltonian does not match size of real part of Hamiltonian.")
        for i, op in enumerate(list(sym_operators) + list(asym_operators)):
            if op.shape != system_sym.shape:
                raise ValueError(f"Size {op.shape} of operator {i + 1} does match size {system_sym.shape} of system Hamiltonian.")
        u0 = np.asarray(u0, dtype=np.float64)
        v0 = np.asarray(v0, dtype=np.float64)
        if u0.shape != v0.shape:
            raise ValueError("Size of the real part of the initial condition does not match the size of the imaginary part.")
        if u0.ndim != 2:
            raise ValueError("only matrix initial conditions are supported (VectorSchrodingerProb is documented as buggy, "
                             "docs/src/problem_setup.md:5-6)")
        if u0.shape[0] != N:
            raise ValueError("Number of levels in initial condition is inconsistent with the size of system Hamiltonian.")
        if len(sym_operators) != len(asym_operators):
            raise ValueError("Number of symmetric operators does not match number of anti-symmetric operators.")
        if guard_subspace_projector.shape != (2 * N, 2 * N):
            raise ValueError("Guard subspace projector size should be twice the size of the complex-valued system.")
        if N_ess_levels > N:
            raise ValueError("Number of essential levels cannot be greater than the total number of levels.")
        if preconditioner_type not in (IdentityPreconditioner, LUPreconditioner, DiagonalHamiltonianPreconditioner):
            raise ValueError("preconditioner_type is not an AbstractQGDPreconditioner.")
        self.system_sym = system_sym
        self.system_asym = system_asym
        self.sym_operators = list(sym_operators)
        self.asym_operators = list(asym_operators)
        self.u0 = np.asfortranarray(u0)
        self.v0 = np.asfortranarray(v0)
        self.guard_subspace_projector = guard_subspace_projector
        self.tf = float(tf)
        self.nsteps = int(nsteps)
        self.N_initial_conditions = u0.shape[1]
        self.N_ess_levels = int(N_ess_levels)
        self.N_tot_levels = int(N)
        self.N_operators = len(sym_operators)
        self.real_system_size = 2 * int(N)
        self.gmres_abstol = float(gmres_abstol)
        self.gmres_reltol = float(gmres_reltol)
        self.preconditioner_type = preconditioner_type

    @classmethod
    def from_hamiltonian(cls, system_hamiltonian, sym_operators, asym_operators, U0, tf, nsteps, N_ess_levels,
                         guard_subspace_projector=None, gmres_abstol=1e-10, gmres_reltol=1e-10,
                         preconditioner_type=IdentityPreconditioner):
        """Outer constructor (src/SchrodingerProb.jl:167-233): K_s = real(H), S_s = imag(H)."""
        H = system_hamiltonian
        Hd = _dense(H)
        if not np.array_equal(Hd, Hd.conj().T):
            raise ValueError("System Hamiltonian is not Hermitian.")
        sparse = sp.issparse(H)
        conv = (lambda M: sp.csc_matrix(np.asarray(_dense(M), dtype=np.float64))) if sparse else (
            lambda M: np.asarray(_dense(M), dtype=np.float64))
        system_sym = conv(Hd.real)
        system_asym = conv(Hd.imag)
        sym_ops = [conv(op) for op in sym_operators]
        asym_ops = [conv(op) for op in asym_operators]
        U0 = np.asarray(U0)
        if U0.ndim == 1:
            U0 = U0[:, None]
        n2 = 2 * Hd.shape[0]
        if guard_subspace_projector is None:
            guard_subspace_projector = np.zeros((n2, n2))
        guard = conv(guard_subspace_projector)
        return cls(system_sym, system_asym, sym_ops, asym_ops, np.real(U0).astype(np.float64),
                   np.imag(U0).astype(np.float64), guard, tf, nsteps, N_ess_levels, gmres_abstol, gmres_reltol,
                   preconditioner_type)

    def copy(self):
        return SchrodingerProb(self.system_sym.copy(), self.system_asym.copy(), [m.copy() for m in self.sym_operators],
                               [m.copy() for m in self.asym_operators], self.u0.copy(), self.v0.copy(),
                               self.guard_subspace_projector.copy(), self.tf, self.nsteps, self.N_ess_levels,
                               self.gmres_abstol, self.gmres_reltol, self.preconditioner_type)


# ---------------------------------------------------------------------------------------------------
# multi_qudit_systems.jl
# ---------------------------------------------------------------------------------------------------
def lowering_operator_subsystem(n: int) -> np.ndarray:
    """:354-359"""
    return np.sqrt(np.diag(np.arange(1, n, dtype=np.float64), k=1))


def lowering_operators_system(subsystem_sizes: Sequence[int]):
    """:364-389 -- kron(I, .., a_i, .., I), first subsystem most significant."""
    ops = []
    for i, n in enumerate(subsystem_sizes):
        mats = [np.eye(s) for s in subsystem_sizes]
        mats[i] = lowering_operator_subsystem(n)
        out = mats[0]
        for M in mats[1:]:
            out = np.kron(out, M)
        ops.append(out)
    return ops


def multi_qudit_hamiltonian_dispersive(subsystem_sizes, transition_freqs, rotation_freqs, kerr_coeffs, sparse_rep=True):
    """:26-58"""
    kerr = np.asarray(kerr_coeffs, dtype=np.float64)
    assert len(transition_freqs) == kerr.shape[0] == kerr.shape[1]
    assert np.array_equal(kerr, kerr.T)
    Q = len(subsystem_sizes)
    n = int(np.prod(subsystem_sizes))
    H = np.zeros((n, n), dtype=np.complex128)
    low = lowering_operators_system(subsystem_sizes)
    for q in range(Q):
        a = low[q]
        H += (transition_freqs[q] - rotation_freqs[q]) * (a.T @ a)
        H -= 0.5 * kerr[q, q] * (a.T @ a.T @ a @ a)
        for p in range(q + 1, Q):
            ap = low[p]
            H -= kerr[p, q] * (ap.T @ ap @ a.T @ a)
    return sp.csc_matrix(H) if sparse_rep else H


def control_ops(subsystem_sizes, sparse_rep=True):
    """:60-71"""
    low = lowering_operators_system(subsystem_sizes)
    sym = [a + a.T for a in low]
    asym = [a - a.T for a in low]
    if sparse_rep:
        return [sp.csc_matrix(m) for m in sym], [sp.csc_matrix(m) for m in asym]
    return sym, asym


def _kron_index(subsystem_sizes, idx):
    k = 0
    for s, i in zip(subsystem_sizes, idx):
        k = k * s + i
    return k


def basis_state(subsystem_sizes, subsystem_indices):
    """:236-259 (bitstring ordered): unit vector at the kron index of |n_0 n_1 ...>."""
    if any(i >= s for i, s in zip(subsystem_indices, subsystem_sizes)):
        raise ValueError(f"Subsystem indices {subsystem_indices} are invalid for subsystem sizes {subsystem_sizes}.")
    v = np.zeros(int(np.prod(subsystem_sizes)))
    v[_kron_index(subsystem_sizes, subsystem_indices)] = 1.0
    return v


def _essential_states(essential_subsystem_sizes):
    """Julia iterates product(reverse(ranges)...) (first range fastest) and reverses each tuple back
    (:262-277): index tuples in subsystem order with the LAST subsystem fastest == itertools.product."""
    return [tuple(t) for t in itertools.product(*[range(e) for e in essential_subsystem_sizes])]


def create_initial_conditions(subsystem_sizes, essential_subsystem_sizes):
    """:255-279 -- column i is the i-th essential basis state (last subsystem fastest)."""
    n = int(np.prod(subsystem_sizes))
    ne = int(np.prod(essential_subsystem_sizes))
    U0 = np.zeros((n, ne), dtype=np.complex128)
    for i, idx in enumerate(_essential_states(essential_subsystem_sizes)):
        U0[:, i] = basis_state(subsystem_sizes, idx)
    return U0


def guard_projector(subsystem_sizes, essential_subsystem_sizes):
    """:316-349.  As written the essential test compares the REVERSED index tuple with the
    un-reversed essential sizes (identical for uniform sizes; kept for fidelity)."""
    n = int(np.prod(subsystem_sizes))
    g = np.zeros(n)
    for i, tup in enumerate(itertools.product(*[range(s) for s in subsystem_sizes])):
        rev_idx = tuple(reversed(tup))  # Julia's `subsystem_indices` before it is reversed back
        if not all(a < b for a, b in zip(rev_idx, essential_subsystem_sizes)):
            g[i] = 1.0
    G = sp.diags(g).tocsc()
    Z = sp.csc_matrix((n, n))
    return sp.bmat([[G, Z], [Z, G]], format="csc")


def create_gate(subsystem_sizes, essential_subsystem_sizes, initial_final_pairs):
    """:391-410 -- initial_final_pairs: iterable of (initial_tuple, final_tuple)."""
    G = create_initial_conditions(subsystem_sizes, essential_subsystem_sizes)
    states = _essential_states(essential_subsystem_sizes)
    for first, second in initial_final_pairs:
        i = states.index(tuple(first))
        G[:, i] = basis_state(subsystem_sizes, tuple(second))
    return G


def DispersiveProblem(subsystem_sizes, essential_subsystem_sizes, transition_freqs, rotation_freqs, kerr_coeffs, tf,
                      nsteps, sparse_rep=True, gmres_abstol=1e-10, gmres_reltol=1e-10,
                      preconditioner_type=DiagonalHamiltonianPreconditioner):
    """:118-165"""
    H = multi_qudit_hamiltonian_dispersive(subsystem_sizes, transition_freqs, rotation_freqs, kerr_coeffs, sparse_rep)
    sym_ops, asym_ops = control_ops(subsystem_sizes)  # (always sparse here, converted to H's type below)
    guard = guard_projector(subsystem_sizes, essential_subsystem_sizes)
    N_ess = int(np.prod(essential_subsystem_sizes))
    U0 = create_initial_conditions(subsystem_sizes, essential_subsystem_sizes)
    return SchrodingerProb.from_hamiltonian(H, sym_ops, asym_ops, U0, tf, nsteps, N_ess, guard,
                                            gmres_abstol=gmres_abstol, gmres_reltol=gmres_reltol,
                                            preconditioner_type=preconditioner_type)


# ---------------------------------------------------------------------------------------------------
# rabi_oscillator.jl / random_problem.jl
# ---------------------------------------------------------------------------------------------------
def construct_rabi_prob(tf=np.pi, gmres_abstol=1e-10, gmres_reltol=1e-10, nsteps=100):
    """rabi_oscillator.jl:7-22"""
    a = np.array([[0.0, 1.0], [0.0, 0.0]])
    return SchrodingerProb.from_hamiltonian(np.zeros((2, 2)), [a + a.T], [a - a.T], np.eye(2), tf, nsteps, 2,
                                            gmres_abstol=gmres_abstol, gmres_reltol=gmres_reltol)


def construct_rand_prob(complex_system_size, N_operators, tf=2.0, nsteps=100, gmres_abstol=1e-10, gmres_reltol=1e-10,
                        seed=0):
    """random_problem.jl:15-35 with numpy's default_rng (Julia's MersenneTwister stream is not reproducible here)."""
    n = complex_system_size

    def rsym(s):
        r = np.random.default_rng(s).random((n, n))
        return r + r.T

    def rasym(s):
        r = np.random.default_rng(s).random((n, n))
        return r - r.T

    rng = np.random.default_rng(seed)
    U0 = rng.random((n, n)) + 1j * rng.random((n, n))
    H = rsym(seed + 2) + 1j * rasym(seed + 3)
    sym_ops = [rsym(seed + 100 + i) for i in range(1, N_operators + 1)]
    asym_ops = [rasym(seed + 200 + i) for i in range(1, N_operators + 1)]
    return SchrodingerProb.from_hamiltonian(H, sym_ops, asym_ops, U0, tf, nsteps, n,
                                            gmres_abstol=gmres_abstol, gmres_reltol=gmres_reltol)


def complex_to_real(x):
    """state_vector_helpers.jl:73-75"""
    x = np.asarray(x)
    return np.asfortranarray(np.concatenate([np.real(x), np.imag(x)], axis=0).astype(np.float64))


def real_to_complex(x):
    """state_vector_helpers.jl:79-88"""
    x = np.asarray(x)
    N = x.shape[0] // 2
    return x[:N] + 1j * x[N:2 * N]
