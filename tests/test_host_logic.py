"""CPU tests of the host-side logic added around the hot path: host-evaluated control tables, Richardson / step-size
helpers, control descriptors of the table path (no GPU, no compute calls into the CUDA library)."""
from math import factorial

import numpy as np
import pytest


def test_sincos_control_tables_closed_form(q):
    """build_control_tables for SinCosControl (src/Controls/sincos_control.jl:1-27): values are d^j/dt^j of
    sin(w t) theta_1 and cos(w t) theta_2 over j!, the table is their derivative w.r.t. theta, layouts as
    include/qgd_b200.h documents for qgd_*_tables."""
    tf, nsteps, m = 2.0, 5, 3
    ctl = [q.SinCosControl(tf, frequency=3.0), q.SinCosControl(tf, frequency=0.5)]
    pcofs = np.array([[0.3, -0.2], [1.1, 0.4], [0.7, 0.9], [-0.5, 0.25]])  # [P=4, B=2]
    cvals, table = q.build_control_tables(ctl, pcofs, tf, nsteps, m)
    assert cvals.shape == (2, m + 1, 2, nsteps + 1, 2) and cvals.flags.f_contiguous
    assert table.shape == (4, m + 1, 2, nsteps + 1) and table.flags.f_contiguous
    for n in range(nsteps + 1):
        t = n * tf / nsteps
        for k, w in enumerate((3.0, 0.5)):
            for j in range(m + 1):
                s = w ** j * np.sin(w * t + j * np.pi / 2) / factorial(j)
                c = w ** j * np.cos(w * t + j * np.pi / 2) / factorial(j)
                for b in range(2):
                    assert cvals[k, j, 0, n, b] == pytest.approx(s * pcofs[2 * k, b], abs=1e-14)
                    assert cvals[k, j, 1, n, b] == pytest.approx(c * pcofs[2 * k + 1, b], abs=1e-14)
                assert table[2 * k, j, 0, n] == pytest.approx(s, abs=1e-14)
                assert table[2 * k + 1, j, 1, n] == pytest.approx(c, abs=1e-14)
                assert table[2 * k, j, 1, n] == 0.0 and table[2 * k + 1, j, 0, n] == 0.0
    # linear controls: the tables contract back to the values
    P = 4
    for b in range(2):
        recon_p = np.einsum("tjn,t->jn", table[:, :, 0, :], pcofs[:, b])
        assert np.allclose(recon_p[:, :], cvals[0, :, 0, :, b] + cvals[1, :, 0, :, b], atol=1e-14)


def test_host_control_descriptor_and_mixing_rules(q):
    c = q.SinCosControl(1.0)
    d, _ = q.controls.control_descriptor(c)
    assert d.type == q._abi.QGD_CONTROL_HOST_TABLE and d.n_amplitudes == 2 and d.n_carriers == 0
    assert q.has_host_controls([q.GRAPEControl(2, 1.0), c]) and not q.has_host_controls(q.GRAPEControl(2, 1.0))
    with pytest.raises(TypeError):  # device families have no host evaluator: a mixed collection cannot fill tables
        q.build_control_tables([q.GRAPEControl(2, 1.0), c], np.zeros(6), 1.0, 4, 1)
    with pytest.raises(TypeError):
        q.controls.control_descriptor(q.CarrierControl(c, [0.0, 1.0]))
    nl = q.HostEvaluatedControl(1, 1.0, lambda t, pc, nd: (np.zeros(nd), np.zeros(nd), np.zeros((nd, 1)), np.zeros((nd, 1))),
                                linear=False)
    with pytest.raises(ValueError):  # nonlinear in pcof: the shared gradient table only serves one control vector
        q.build_control_tables([nl], np.zeros((1, 2)), 1.0, 2, 1)


def test_richardson_helpers(q):
    """richardson_extrap_sol / _rel_err (src/Tests/test_convergence.jl:238-250) on a model error expansion."""
    exact = np.array([1.0, -2.0, 0.5])
    order, h = 4, 0.1
    c = np.array([0.3, 0.1, -0.2])
    Ah, A2h = exact + c * h ** order, exact + c * (2 * h) ** order
    assert np.allclose(q.richardson_extrap_sol(Ah, A2h, order), exact, atol=1e-15)
    assert q.richardson_extrap_rel_err(Ah, A2h, order) == pytest.approx(np.linalg.norm(c * h ** order) / np.linalg.norm(exact))


def test_estimate_n_timesteps(q):
    """get_shortest_period / estimate_N_timesteps (src/calculate_timestep.jl:17-45) on a diagonal Hamiltonian."""
    H = np.diag([0.0, 2.0, -5.0]).astype(complex)
    prob = q.SchrodingerProb.from_hamiltonian(H, [np.zeros((3, 3))], [np.zeros((3, 3))], np.eye(3, dtype=complex), 10.0, 10, 3)
    assert q.get_shortest_period(prob, [0.0]) == pytest.approx(2 * np.pi / 5.0)
    assert q.estimate_N_timesteps(prob, [0.0], 40) == int(np.ceil(10.0 / (2 * np.pi / 5.0) * 40))
    # the control amplitude enters through max |eig| of the full Hamiltonian
    K = np.zeros((3, 3)); K[0, 1] = K[1, 0] = 1.0
    prob2 = q.SchrodingerProb.from_hamiltonian(H, [K], [np.zeros((3, 3))], np.eye(3, dtype=complex), 10.0, 10, 3)
    assert q.get_shortest_period(prob2, [100.0]) < q.get_shortest_period(prob2, [0.0])


def test_forced_and_table_entry_points_reject_bad_arguments_without_gpu(q):
    """No GPU here: handle creation must fail loudly (no CPU fallback) also for host-evaluated controls."""
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    prob = q.construct_rand_prob(4, 1, tf=1.0, nsteps=4)
    with pytest.raises(q.QGDError):
        q.Handle(prob, [q.SinCosControl(1.0)])
