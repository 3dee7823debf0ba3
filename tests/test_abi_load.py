"""CPU-side checks of the drop-in boundary: the C-ABI library loads, exports every symbol
include/qgd_b200.h declares, and fails loudly (no CPU fallback) without a GPU."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_functions():
    txt = open(os.path.join(ROOT, "include", "qgd_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(qgd_[a-z0-9_]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol(q):
    if not os.path.exists(q.backend.LIB_PATH):
        q.backend.build()
    lib = ctypes.CDLL(q.backend.LIB_PATH)
    declared = _header_functions()
    assert len(declared) >= 15
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/qgd_b200.h but not exported"
    assert set(q.backend.EXPORTS) <= set(declared)


def test_struct_layout_matches_header(q):
    # sizes of the ABI structs as the C compiler sees them (LP64): guards the ctypes mirror
    A = q._abi
    assert ctypes.sizeof(A.qgd_matrix_t) == 64
    assert ctypes.sizeof(A.qgd_control_t) == 64
    assert ctypes.sizeof(A.qgd_problem_t) == 32 + 2 * 64 + 4 * 8 + 64 + 8 + 8 + 16 + 8 + 8
    assert ctypes.sizeof(A.qgd_stats_t) == 64


def test_n_coeff_helpers(q):
    prob, controls, pcof, target, order = q.configs.cnot3(nsteps=4, tf=4.0, subsystem_sizes=(2, 2, 2))
    pk = q._abi.ProblemPack(prob, controls)
    lib = q.backend.lib()
    assert lib.qgd_problem_n_coeff(pk.ref()) == len(pcof) == 180


def test_no_cpu_fallback_without_gpu(q):
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    prob, controls, pcof, target, order = q.configs.cnot2(nsteps=4, tf=4.0)
    with pytest.raises(q.QGDError) as ei:
        q.Handle(prob, controls)
    assert ei.value.code == -2  # QGD_ECUDA
    with pytest.raises(q.QGDError):
        q.discrete_adjoint(prob, controls, pcof, target, order=order)


def test_plain_c_program_links_and_fails_loudly_without_gpu(q):
    """tests/abi_c/abi_smoke.c links against libqgd_b200.so with gcc (no Python in between); on this CPU box qgd_create
    returns QGD_ECUDA and the program says so -- the GPU suite runs the same binary through create -> eval -> destroy."""
    import subprocess

    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present: covered by tests/test_gpu_round2.py::test_c_program_drives_the_abi")
    if not os.path.exists(q.backend.LIB_PATH):
        q.backend.build()
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "tests", "abi_c")])
    r = subprocess.run([os.path.join(ROOT, "tests", "abi_c", "abi_smoke")], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and r.stdout.startswith("OK (no device"), (r.stdout, r.stderr)


def test_library_reads_no_environment_variables(q):
    """Behaviour switches are ABI calls (qgd_set_option), not getenv."""
    csrc = os.path.join(ROOT, "quantumgatedesign.jl_b200", "csrc")
    for fn in os.listdir(csrc):
        if fn.endswith((".cu", ".cuh", ".h")):
            assert "getenv" not in open(os.path.join(csrc, fn)).read(), fn


def test_handle_cache_key_is_content_based(q):
    """the api's handle cache was keyed on id(prob) / id(control); CPython reuses ids of freed
    objects, so a temporary problem could inherit a dead problem's device operators.  The key is now a content
    fingerprint: distinct problems -> distinct keys, equal content -> equal key, mutable knobs excluded."""
    keys = set()
    for i in range(40):
        prob = q.construct_rand_prob(4, 1, tf=1.0, nsteps=6, seed=1000 + i)
        keys.add(q.backend.problem_key(prob, q.GRAPEControl(3, prob.tf)))
        del prob
    assert len(keys) == 40
    p1 = q.construct_rand_prob(4, 1, tf=1.0, nsteps=6, seed=5)
    p2 = q.construct_rand_prob(4, 1, tf=1.0, nsteps=6, seed=5)
    c = q.GRAPEControl(3, 1.0)
    assert q.backend.problem_key(p1, c) == q.backend.problem_key(p2, q.GRAPEControl(3, 1.0))
    assert q.backend.problem_key(p1, c) != q.backend.problem_key(p1, q.GRAPEControl(4, 1.0))
    assert q.backend.problem_key(p1, c) != q.backend.problem_key(p1, q.CarrierControl(q.GRAPEControl(3, 1.0), [0.0, 1.0]))
    assert q.backend.problem_key(p1, q.CarrierControl(c, [0.0, 1.0])) != q.backend.problem_key(p1, q.CarrierControl(c, [0.0, 2.0]))
    k = q.backend.problem_key(p1, c)
    p1.nsteps = 12; p1.gmres_abstol = 1e-13  # mutable knobs of the reference's SchrodingerProb: re-synced per call, not part of the key
    assert q.backend.problem_key(p1, c) == k


def test_host_validation_mirrors_reference(q):
    a = np.array([[0.0, 1.0], [0.0, 0.0]])
    with pytest.raises(ValueError, match="not symmetric"):
        q.SchrodingerProb(np.array([[0, 1.0], [0, 0]]), np.zeros((2, 2)), [a + a.T], [a - a.T], np.eye(2), np.zeros((2, 2)),
                          np.zeros((4, 4)), 1.0, 10, 2)
    with pytest.raises(ValueError, match="anti-symmetric"):
        q.SchrodingerProb(np.zeros((2, 2)), np.zeros((2, 2)), [a + a.T], [a + a.T], np.eye(2), np.zeros((2, 2)),
                          np.zeros((4, 4)), 1.0, 10, 2)
    with pytest.raises(ValueError, match="Hermitian"):
        q.SchrodingerProb.from_hamiltonian(np.array([[0, 1j], [1j, 0]]), [a + a.T], [a - a.T], np.eye(2), 1.0, 10, 2)
    with pytest.raises(ValueError, match="essential"):
        q.SchrodingerProb.from_hamiltonian(np.zeros((2, 2)), [a + a.T], [a - a.T], np.eye(2), 1.0, 10, 3)
    with pytest.raises(ValueError, match="D1"):
        q.BSpline2Control(2, 1.0)


def test_problem_constructors(q):
    """DispersiveProblem / guard_projector / create_gate shapes (multi_qudit_systems.jl docstring examples)."""
    G = q.guard_projector([3], [2]).toarray()
    assert np.array_equal(np.diag(G), [0, 0, 1, 0, 0, 1])
    G = q.guard_projector([2, 2], [2, 1]).toarray()
    assert np.array_equal(np.diag(G), [0, 0, 1, 1, 0, 0, 1, 1])
    prob, controls, pcof, target, order = q.configs.cnot3(nsteps=4, tf=4.0)
    assert prob.N_tot_levels == 64 and prob.N_initial_conditions == 8 and prob.N_ess_levels == 8
    assert prob.system_sym.nnz <= 64 and prob.system_asym.nnz == 0
    assert all(op.nnz == 96 for op in prob.sym_operators)
    assert np.allclose(np.abs(target).sum(axis=0), 1.0) and target.shape == (64, 8)
    # CNOT on the last two essential qubits: |a10> <-> |a11>
    U0 = q.create_initial_conditions((4, 4, 4), (2, 2, 2))
    assert np.array_equal(target[:, 2], U0[:, 3]) and np.array_equal(target[:, 3], U0[:, 2])
    assert np.array_equal(target[:, 0], U0[:, 0]) and np.array_equal(target[:, 6], U0[:, 7])
