"""Oracle control families vs independent references (mirrors the reference's
test/ControlFunctionTests/test_control_derivatives.jl and test_control_gradients.jl)."""
import numpy as np
import pytest
from scipy.interpolate import BSpline


def _fbs_knots(degree, n_basis, tf):
    order = degree + 1
    n_knots = n_basis + order
    nd = n_knots - 2 * (order - 1)
    return np.concatenate([np.zeros(order - 1), np.linspace(0, 1, nd), np.ones(order - 1)]) * tf


@pytest.mark.parametrize("degree", [2, 4, 8, 16])
def test_pppack_port_matches_scipy(q, O, degree):
    """FortranBSplineControl values and derivatives r=0..4 vs scipy.interpolate.BSpline."""
    tf, n_basis = 3.7, 20
    ctl = q.FortranBSplineControl(degree, n_basis, tf)
    prob = q.construct_rabi_prob(tf=tf)
    rng = np.random.default_rng(degree)
    pcof = rng.standard_normal(ctl.N_coeff)
    kn = _fbs_knots(degree, n_basis, tf)
    sp_p = BSpline(kn, pcof[:n_basis], degree)
    sp_q = BSpline(kn, pcof[n_basis:], degree)
    ts = np.concatenate([[0.0], rng.random(40) * tf, [tf * (1 - 1e-12)]])
    for t in ts:
        for r in range(0, 5):
            pv, qv, _, _ = O.eval_pq_derivative(prob, ctl, 0, pcof, t, r, want_grad=False)
            ep = 0.0 if r > degree else float(sp_p(t, nu=r))
            eq = 0.0 if r > degree else float(sp_q(t, nu=r))
            scale = max(1.0, abs(ep), abs(eq))
            assert abs(pv - ep) <= 2e-11 * scale * (n_basis / tf) ** r
            assert abs(qv - eq) <= 2e-11 * scale * (n_basis / tf) ** r


def test_fortran_bspline_high_derivative_is_zero(q, O):
    """pppack clamps mhigh = min(nderiv, k) (bsplvd.f:44): derivatives of order >= degree+1 vanish."""
    ctl = q.FortranBSplineControl(2, 10, 1.0)
    prob = q.construct_rabi_prob(tf=1.0)
    pcof = np.arange(1.0, 21.0)
    for r in (3, 4, 5):
        pv, qv, gp, gq = O.eval_pq_derivative(prob, ctl, 0, pcof, 0.37, r)
        assert pv == 0.0 and qv == 0.0 and not gp.any() and not gq.any()


def _controls(q, tf):
    fbs = q.FortranBSplineControl(8, 10, tf)
    return {
        "grape": q.GRAPEControl(10, tf),
        "bspline2": q.BSpline2Control(10, tf),
        "fbs2": q.FortranBSplineControl(2, 10, tf),
        "fbs6": q.FortranBSplineControl(6, 10, tf),
        "carrier_fbs8": q.CarrierControl(fbs, [-10, -1, 0, 1, 10]),
        "carrier_bs2": q.CarrierControl(q.BSpline2Control(10, tf), [0.0, 0.3, -1.7]),
    }


@pytest.mark.parametrize("name", ["grape", "bspline2", "fbs2", "fbs6", "carrier_fbs8", "carrier_bs2"])
def test_time_derivatives_vs_central_difference(q, O, name):
    """eval_{p,q}_derivative(order) vs central difference of order-1 (test_control_derivatives.jl:14-90):
    >= 95% of 1000 points within 50*(1e-15)^(2/3)-scaled tolerance."""
    tf = 2.0
    ctl = _controls(q, tf)[name]
    prob = q.construct_rabi_prob(tf=tf)
    rng = np.random.default_rng(3)
    pcof = rng.random(ctl.N_coeff)
    h = 1e-5
    ts = np.linspace(0.01, tf - 0.01, 1000)
    for order in (1, 2, 3, 4):
        ok = 0
        for t in ts:
            pv, qv, _, _ = O.eval_pq_derivative(prob, ctl, 0, pcof, t, order, want_grad=False)
            pr_, qr_, _, _ = O.eval_pq_derivative(prob, ctl, 0, pcof, t + h, order - 1, want_grad=False)
            pl_, ql_, _, _ = O.eval_pq_derivative(prob, ctl, 0, pcof, t - h, order - 1, want_grad=False)
            fp, fq = (pr_ - pl_) / (2 * h), (qr_ - ql_) / (2 * h)
            tol = 1e-5 * max(1.0, abs(pv), abs(qv), abs(pr_) / h * 1e-10)
            ok += (abs(pv - fp) <= tol) and (abs(qv - fq) <= tol)
        assert ok >= 0.95 * len(ts), (name, order, ok)


@pytest.mark.parametrize("name", ["grape", "bspline2", "fbs2", "fbs6", "carrier_fbs8", "carrier_bs2"])
def test_parameter_gradients_are_exact_linear_maps(q, O, name):
    """Every control on the path is linear in pcof: <grad p^(r), pcof> == p^(r) and the gradient is
    pcof-independent (test_control_gradients.jl:9-113 checks the same thing by FD in pcof)."""
    tf = 2.0
    ctl = _controls(q, tf)[name]
    prob = q.construct_rabi_prob(tf=tf)
    rng = np.random.default_rng(5)
    pcof = rng.standard_normal(ctl.N_coeff)
    other = rng.standard_normal(ctl.N_coeff)
    for t in rng.random(25) * tf:
        for r in range(0, 5):
            pv, qv, gp, gq = O.eval_pq_derivative(prob, ctl, 0, pcof, t, r)
            _, _, gp2, gq2 = O.eval_pq_derivative(prob, ctl, 0, other, t, r)
            s = max(1.0, np.abs(gp).max(), np.abs(gq).max())
            assert abs(gp @ pcof - pv) <= 1e-12 * s * ctl.N_coeff
            assert abs(gq @ pcof - qv) <= 1e-12 * s * ctl.N_coeff
            assert np.array_equal(gp, gp2) and np.array_equal(gq, gq2)


def test_bspline2_partition_of_unity(q, O):
    """Quadratic B-splines with all coefficients 1 sum to 1 inside [0, tf] (Juqbox-style basis)."""
    tf = 5.0
    ctl = q.BSpline2Control(10, tf)
    prob = q.construct_rabi_prob(tf=tf)
    pcof = np.ones(ctl.N_coeff)
    for t in np.linspace(0, tf, 57):
        pv, qv, _, _ = O.eval_pq_derivative(prob, ctl, 0, pcof, t, 0, want_grad=False)
        assert abs(pv - 1) < 1e-13 and abs(qv - 1) < 1e-13


def test_fill_mat_is_taylor_scaled(q, O):
    """fill_p_mat!/fill_q_mat! entry (1+j,k) = p_k^(j)(t)/j! (Control.jl:99-149)."""
    import math

    tf = 2.0
    ctls = [q.FortranBSplineControl(6, 12, tf), q.CarrierControl(q.BSpline2Control(8, tf), [0.0, 1.1])]
    a = np.array([[0.0, 1.0], [0.0, 0.0]])
    prob = q.SchrodingerProb.from_hamiltonian(np.zeros((2, 2)), [a + a.T] * 2, [a - a.T] * 2, np.eye(2), tf, 10, 2)
    rng = np.random.default_rng(11)
    pcof = rng.standard_normal(q.get_number_of_control_parameters(ctls))
    sl = q.control_slices(ctls)
    for t in (0.0, 0.3, 1.234, tf):
        P, Q = O.fill_pq_mat(prob, ctls, pcof, t, 5)
        for k in range(2):
            for j in range(5):
                pv, qv, _, _ = O.eval_pq_derivative(prob, ctls, k, pcof[sl[k][0]:sl[k][1]], t, j, want_grad=False)
                assert abs(P[j, k] - pv / math.factorial(j)) <= 1e-13 * max(1, abs(pv))
                assert abs(Q[j, k] - qv / math.factorial(j)) <= 1e-13 * max(1, abs(qv))


def test_grape_domain_error(q, O):
    ctl = q.GRAPEControl(4, 1.0)
    prob = q.construct_rabi_prob(tf=1.0)
    with pytest.raises(RuntimeError):
        O.eval_pq_derivative(prob, ctl, 0, np.ones(8), 1.5, 0)
