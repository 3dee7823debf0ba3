"""Golden fixtures (tests/golden/*.npz, written by tests/golden/make_golden.py from the CPU oracle).

CPU: the oracle still reproduces them (the restatement has not drifted) and the input generators still produce
the same inputs.  GPU (-m gpu): the CUDA path through the C ABI reproduces them WITHOUT the oracle in the loop.
Tolerance: 1e-10 relative on infidelity, gradient and final state (north_star), equal GMRES iteration counts."""
import importlib.util
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
spec = importlib.util.spec_from_file_location("make_golden", os.path.join(HERE, "golden", "make_golden.py"))
mg = importlib.util.module_from_spec(spec)
spec.loader.exec_module(mg)

NAMES = ["cnot2_o4", "cnot3_333_o8", "cnot3_444_o8_short", "dense_o10"]
RTOL = 1e-10


def rel(a, b):
    a, b = np.asarray(a), np.asarray(b)
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-300)


def load(name):
    return np.load(os.path.join(HERE, "golden", name + ".npz"))


@pytest.mark.parametrize("name", NAMES)
def test_inputs_are_reproducible(q, name):
    prob, controls, pcof, target, order = mg.cases(q)[name]
    g = load(name)
    assert int(g["order"]) == order
    assert str(g["digest"]) == mg.input_digest(q, prob, controls, pcof, target)


@pytest.mark.parametrize("name", NAMES)
def test_oracle_reproduces_golden(q, O, name):
    prob, controls, pcof, target, order = mg.cases(q)[name]
    g = load(name)
    ref = O.discrete_adjoint(prob, controls, pcof, target, order=order)
    assert rel(ref["grad"], g["grad"]) < 1e-12
    assert abs(ref["infidelity"] - float(g["infidelity"])) <= 1e-12 * abs(float(g["infidelity"]))
    assert rel(ref["history"][:, 0, -1, :], g["final_state"]) < 1e-12
    assert np.array_equal(ref["iters_fwd"], g["iters_fwd"]) and np.array_equal(ref["iters_adj"], g["iters_adj"])


@pytest.mark.gpu
@pytest.mark.parametrize("name", NAMES)
def test_cuda_path_reproduces_golden(q, name):
    prob, controls, pcof, target, order = mg.cases(q)[name]
    g = load(name)
    h = q.Handle(prob, controls)
    out = h.discrete_adjoint(pcof, q.complex_to_real(target), order=order, want_history=True, want_iters=True)
    h.close()
    assert rel(out["grad"][:, 0], g["grad"]) < RTOL
    assert abs(out["infidelity"][0] - float(g["infidelity"])) <= RTOL * abs(float(g["infidelity"]))
    assert abs(out["guard_penalty"][0] - float(g["guard_penalty"])) <= RTOL * max(abs(float(g["guard_penalty"])), 1e-300)
    assert rel(out["history"][:, 0, -1, :, 0], g["final_state"]) < RTOL
    assert np.array_equal(out["iters_fwd"][:, :, 0], g["iters_fwd"])
    assert np.array_equal(out["iters_adj"][:, :, 0], g["iters_adj"])
    assert np.array_equal(out["iters_term"][:, 0], g["iters_term"])
