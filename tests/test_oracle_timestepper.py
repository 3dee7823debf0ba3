"""Oracle time stepper vs known answers."""
import math

import numpy as np
import pytest
import scipy.linalg as sla


def _rand_ops(n, seed):
    rng = np.random.default_rng(seed)
    r = rng.random((n, n)); Ks = r + r.T
    r = rng.random((n, n)); Ss = r - r.T
    r = rng.random((n, n)); Kc = r + r.T
    r = rng.random((n, n)); Sc = r - r.T
    return Ks, Ss, Kc, Sc


def _A(K, S):
    return np.block([[S, K], [-K, S]])


def test_derivatives_closed_form(q, O):
    """compute_derivatives!/compute_adjoint_derivatives! vs the explicit matrix formulas of the reference's
    test/hardcoded_derivatives.jl:77-130: w1=Aw, w2=(A'+A^2)w/2!, w3=(A''+2A'A+AA'+A^3)w/3!, w4=(...)w/4!"""
    n = 3
    Ks, Ss, Kc, Sc = _rand_ops(n, 0)
    prob = q.SchrodingerProb.from_hamiltonian(Ks + 1j * Ss, [Kc], [Sc], np.eye(n), 1.0, 1, n)
    ctl = q.GRAPEControl(1, 1.0)
    rng = np.random.default_rng(1)
    m = 4
    cre = rng.standard_normal((m + 1, 1)); cim = rng.standard_normal((m + 1, 1))
    # A^(j) = j! * (c_re[j] * [0 Kc; -Kc 0] + c_im[j] * [Sc 0; 0 Sc]) (+ drift for j=0)
    Z = np.zeros((n, n))
    dA = [math.factorial(j) * (cre[j, 0] * _A(Kc, Z) + cim[j, 0] * _A(Z, Sc)) for j in range(m + 1)]
    A0 = dA[0] + _A(Ks, Ss)
    A1, A2, A3 = dA[1], dA[2], dA[3]
    M = [np.eye(2 * n), A0, (A1 + A0 @ A0) / 2,
         (A2 + 2 * A1 @ A0 + A0 @ A1 + A0 @ A0 @ A0) / 6,
         (A3 + 3 * A2 @ A0 + 3 * A1 @ A1 + 3 * A1 @ A0 @ A0 + A0 @ A2 + 2 * A0 @ A1 @ A0 + A0 @ A0 @ A1
          + np.linalg.matrix_power(A0, 4)) / 24]
    w = rng.standard_normal(2 * n)
    uv = np.zeros((2 * n, m + 1), order="F"); uv[:, 0] = w
    out = O.compute_derivatives(prob, ctl, uv, 2 * m, cre, cim, adjoint=False)
    outT = O.compute_derivatives(prob, ctl, uv, 2 * m, cre, cim, adjoint=True)
    for j in range(m + 1):
        ref = M[j] @ w
        refT = M[j].T @ w
        assert np.allclose(out[:, j], ref, rtol=1e-13, atol=1e-13 * np.abs(ref).max())
        assert np.allclose(outT[:, j], refT, rtol=1e-13, atol=1e-13 * np.abs(refT).max())


def test_adjoint_operator_is_transpose(q, O):
    """LHSHolderAdjoint == transpose of LHSHolder as matrices (A_j^T = -A_j), any order."""
    n = 3
    Ks, Ss, Kc, Sc = _rand_ops(n, 4)
    prob = q.SchrodingerProb.from_hamiltonian(Ks + 1j * Ss, [Kc], [Sc], np.eye(n), 0.8, 4, n)
    ctl = q.GRAPEControl(1, 0.8)
    rng = np.random.default_rng(2)
    for order in (2, 6, 10):
        m = order // 2
        cre = rng.standard_normal((m + 1, 1)); cim = rng.standard_normal((m + 1, 1))
        I = np.eye(2 * n)
        L = np.column_stack([O.apply_step_operator(prob, ctl, I[:, i], order, cre, cim, False, True) for i in range(2 * n)])
        LT = np.column_stack([O.apply_step_operator(prob, ctl, I[:, i], order, cre, cim, True, True) for i in range(2 * n)])
        assert np.allclose(LT, L.T, rtol=1e-12, atol=1e-12 * np.abs(L).max())


def test_rabi_pi_pulse_gives_x_gate(q, O):
    """Constant p=0.5, q=0, tf=pi produces the SWAP/X gate up to phase
    (src/ProblemConstructors/rabi_oscillator.jl:1-6): infidelity -> 0 at the method's order."""
    prob = q.construct_rabi_prob(tf=np.pi, gmres_abstol=1e-15, gmres_reltol=1e-15, nsteps=20)
    ctl = q.GRAPEControl(1, prob.tf)
    pcof = np.array([0.5, 0.0])
    target = q.complex_to_real(np.array([[0, 1], [1, 0]], dtype=complex))
    prev = None
    for order in (2, 4, 6, 8):
        h, _ = O.eval_forward(prob, ctl, pcof, order=order)
        inf = O.infidelity_real(h[:, 0, -1, :], target, 2)
        assert inf < (1e-4 if order == 2 else 1e-9)
        if prev is not None:
            assert inf <= prev + 1e-15
        prev = inf
    assert prev < 1e-13


@pytest.mark.parametrize("order", [2, 4, 6, 8])
def test_forward_convergence_order(q, O, order):
    """Step-halving: error vs the exact propagator for a time-independent random Hamiltonian decays
    with slope = order +- 0.5 (test/ConvergenceTests/forward_convergence.jl:55-65)."""
    n = 4
    Ks, Ss, Kc, Sc = _rand_ops(n, 9)
    tf = 1.0
    ctl = q.GRAPEControl(1, tf)
    pcof = np.array([0.3, -0.2])
    A = _A(Ks + 0.3 * Kc, Ss - 0.2 * Sc)
    exact = sla.expm(A * tf) @ np.vstack([np.eye(n), np.zeros((n, n))])
    errs, hs = [], []
    for nsteps in (4, 8, 16, 32):
        prob = q.SchrodingerProb.from_hamiltonian(Ks + 1j * Ss, [Kc], [Sc], np.eye(n), tf, nsteps, n,
                                                  gmres_abstol=1e-15, gmres_reltol=1e-15)
        h, _ = O.eval_forward(prob, ctl, pcof, order=order)
        errs.append(np.abs(h[:, 0, -1, :] - exact).max())
        hs.append(tf / nsteps)
    errs = np.array(errs); hs = np.array(hs)
    good = errs > 1e-12  # above roundoff
    assert good.sum() >= 2
    slope = np.polyfit(np.log(hs[good]), np.log(errs[good]), 1)[0]
    assert abs(slope - order) < 0.6, (order, slope, errs)


def test_time_dependent_convergence_with_bspline_carrier(q, O):
    """Order-6 vs order-12 solutions with smooth (degree 16) carrier controls agree to high accuracy and the
    order-6 error drops ~2^6 per halving."""
    prob0, controls, pcof, target, _ = q.configs.dense_random(N=4, Nc=1, nsteps=8, degree=16, n_basis=20,
                                                              carriers=(0.0, 1.0), dt_norm=0.25)
    def final(nsteps, order):
        p = prob0.copy(); p.nsteps = nsteps; p.gmres_abstol = p.gmres_reltol = 1e-15
        h, _ = O.eval_forward(p, controls, pcof, order=order)
        return h[:, 0, -1, :]
    ref = final(64, 12)
    e1 = np.abs(final(8, 6) - ref).max()
    e2 = np.abs(final(16, 6) - ref).max()
    assert e2 < e1 and 5.0 < np.log2(e1 / e2) < 7.2, (e1, e2)


def test_save_every_nsteps(q, O):
    prob, controls, pcof, target, order = q.configs.cnot2(nsteps=12, tf=12.0, gmres_tol=1e-13)
    h1, _ = O.eval_forward(prob, controls, pcof, order=order, saveEveryNsteps=1)
    h3, _ = O.eval_forward(prob, controls, pcof, order=order, saveEveryNsteps=3)
    assert h3.shape[2] == 5
    assert np.array_equal(h3[:, :, :, :], h1[:, :, ::3, :])


def test_gmres_matches_dense_solve(O):
    """IterativeSolvers-style GMRES (restarted, MGS, null-vector residual) on a dense system."""
    rng = np.random.default_rng(0)
    n = 30
    A = np.eye(n) + 0.3 * rng.standard_normal((n, n)) / np.sqrt(n)
    b = rng.standard_normal(n)
    x, it = O.gmres_dense(A, b, abstol=1e-13, reltol=0.0, restart=n, maxiter=n)
    assert np.allclose(A @ x, b, atol=1e-11) and it <= n
    x2, it2 = O.gmres_dense(A, b, abstol=1e-13, reltol=0.0, restart=5, maxiter=200)  # with restarts
    assert np.allclose(A @ x2, b, atol=1e-10) and it2 >= it
    x3, it3 = O.gmres_dense(A, b, x0=x, abstol=1e-9, reltol=0.0)  # converged initial guess: zero iterations
    assert it3 == 0 and np.array_equal(x3, x)


def test_diagonal_preconditioner_exact_for_order2(q, O):
    """For order 2 the no-control LHS is I - dt/2 A_d; with a diagonal drift the Diagonal preconditioner
    inverts it exactly (preconditioners.jl:57-62)."""
    prob, controls, pcof, target, _ = q.configs.cnot3(nsteps=10, tf=10.0, subsystem_sizes=(3, 3, 3))
    n2 = prob.real_system_size
    rng = np.random.default_rng(0)
    x = rng.standard_normal(n2)
    m = 1
    zero = np.zeros((m + 1, prob.N_operators))
    for adjoint in (False, True):
        Lx = O.apply_step_operator(prob, controls, x, 2, zero, zero, adjoint, True)
        back = O.apply_preconditioner(prob, controls, Lx, 2, adjoint)
        assert np.allclose(back, x, rtol=1e-13, atol=1e-13)
