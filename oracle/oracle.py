"""ctypes front end of the CPU ORACLE (oracle/qgd_oracle.cpp).

TEST INFRASTRUCTURE ONLY -- PARITY UNPINNED (see the header of qgd_oracle.cpp): may be imported by
tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs, never by the
product package.  It shares only the data-marshalling structs (`_abi.ProblemPack`) with the product.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def _pkg():
    import sys

    sys.path.insert(0, os.path.dirname(_HERE))
    try:
        from __graft_entry__ import load_package
    finally:
        sys.path.pop(0)
    return load_package()


def build(force: bool = False) -> str:
    so = os.path.join(_HERE, "libqgd_oracle.so")
    src = os.path.join(_HERE, "qgd_oracle.cpp")
    hdr = os.path.join(os.path.dirname(_HERE), "include", "qgd_b200.h")
    stale = (not os.path.exists(so)) or any(
        os.path.exists(f) and os.path.getmtime(f) > os.path.getmtime(so) for f in (src, hdr)
    )
    if force or stale:
        subprocess.check_call(["make", "-C", _HERE, "-B" if force else "-s"], stdout=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(build())
        _LIB.qgdo_last_error.restype = C.c_char_p
    return _LIB


def _check(rc):
    if rc != 0:
        raise RuntimeError("oracle: " + lib().qgdo_last_error().decode())


def _d(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _i(a):
    return a.ctypes.data_as(C.POINTER(C.c_int64))


def _pack(prob, controls):
    return _pkg()._abi.ProblemPack(prob, controls)


def eval_forward(prob, controls, pcof, order=2, saveEveryNsteps=1, forcing=None, nthreads=None):
    """-> (history [2N,1+m,1+nsteps/save,nic], gmres_iters [nsteps,nic])"""
    pk = _pack(prob, controls)
    m = order // 2
    nslots = 1 + prob.nsteps // saveEveryNsteps
    hist = np.zeros((prob.real_system_size, 1 + m, nslots, prob.N_initial_conditions), order="F")
    iters = np.zeros((prob.nsteps, prob.N_initial_conditions), dtype=np.int64, order="F")
    pc = np.ascontiguousarray(pcof, dtype=np.float64)
    f = None
    if forcing is not None:
        f = np.asfortranarray(forcing, dtype=np.float64)
    nt = nthreads or min(prob.N_initial_conditions, os.cpu_count() or 1)
    _check(lib().qgdo_eval_forward(pk.ref(), _d(pc), C.c_int(order), C.c_int64(saveEveryNsteps),
                                   _d(f) if f is not None else None, _d(hist), _i(iters), C.c_int(nt)))
    return hist, iters


def discrete_adjoint(prob, controls, pcof, target, order=2, nthreads=None, history=None):
    """target: complex [N,nic] (or real [N,nic]) as in the reference.  -> dict"""
    pkg = _pkg()
    pk = _pack(prob, controls)
    m = order // 2
    n2, nic, Nt = prob.real_system_size, prob.N_initial_conditions, prob.nsteps + 1
    tgt = pkg.complex_to_real(target)
    pc = np.ascontiguousarray(pcof, dtype=np.float64)
    grad = np.zeros(pk.n_coeff)
    pre = history is not None
    hist = np.asfortranarray(history).copy(order="F") if pre else np.zeros((n2, 1 + m, Nt, nic), order="F")
    lam = np.zeros((n2, 1 + m, Nt, nic), order="F")
    forc = np.zeros((n2, Nt, nic), order="F")
    it_f = np.zeros((prob.nsteps, nic), dtype=np.int64, order="F")
    it_a = np.zeros((prob.nsteps, nic), dtype=np.int64, order="F")
    it_t = np.zeros(nic, dtype=np.int64)
    nt = nthreads or min(nic, os.cpu_count() or 1)
    _check(lib().qgdo_discrete_adjoint(pk.ref(), _d(pc), _d(tgt), C.c_int(order), C.c_int(1 if pre else 0), _d(grad),
                                       _d(hist), _d(lam), _d(forc), _i(it_f), _i(it_a), _i(it_t), C.c_int(nt)))
    final = np.asfortranarray(hist[:, 0, -1, :])
    return dict(grad=grad, history=hist, lambda_history=lam, adjoint_forcing=forc, iters_fwd=it_f, iters_adj=it_a,
                iters_term=it_t, final_state=final,
                infidelity=infidelity_real(final, tgt, prob.N_ess_levels),
                guard_penalty=guard_penalty_real(prob, controls, hist, order))


def eval_grad_forced(prob, controls, pcof, target, order=2, nthreads=None):
    pkg = _pkg()
    pk = _pack(prob, controls)
    tgt = pkg.complex_to_real(target)
    pc = np.ascontiguousarray(pcof, dtype=np.float64)
    grad = np.zeros(pk.n_coeff)
    nt = nthreads or min(prob.N_initial_conditions, os.cpu_count() or 1)
    _check(lib().qgdo_eval_grad_forced(pk.ref(), _d(pc), _d(tgt), C.c_int(order), _d(grad), C.c_int(nt)))
    return grad


def infidelity_real(psi, target, N_ess):
    psi = np.asfortranarray(psi, dtype=np.float64)
    target = np.asfortranarray(target, dtype=np.float64)
    if psi.ndim == 1:
        psi, target = psi[:, None], target[:, None]
    out = C.c_double()
    _check(lib().qgdo_infidelity_real(_d(psi), _d(target), C.c_int64(psi.shape[0] // 2), C.c_int64(psi.shape[1]),
                                      C.c_int64(N_ess), C.byref(out)))
    return out.value


def guard_penalty_real(prob, controls, history, order):
    pk = _pack(prob, controls)
    h = np.asfortranarray(history, dtype=np.float64)
    out = C.c_double()
    _check(lib().qgdo_guard_penalty_real(pk.ref(), _d(h), C.c_int(order), C.byref(out)))
    return out.value


def objective(prob, controls, pcof, target, order=2, nthreads=None):
    """infidelity + guard penalty, what eval_grad_finite_difference differentiates
    (src/eval_grad_finite_difference.jl:42-66)."""
    pkg = _pkg()
    hist, _ = eval_forward(prob, controls, pcof, order=order, nthreads=nthreads)
    tgt = pkg.complex_to_real(target)
    return infidelity_real(hist[:, 0, -1, :], tgt, prob.N_ess_levels) + guard_penalty_real(prob, controls, hist, order)


def eval_grad_finite_difference(prob, controls, pcof, target, order=2, dpcof=1e-5, nthreads=None):
    """src/eval_grad_finite_difference.jl:16-72"""
    pcof = np.asarray(pcof, dtype=np.float64)
    g = np.zeros_like(pcof)
    for i in range(pcof.size):
        r = pcof.copy(); r[i] += dpcof
        l = pcof.copy(); l[i] -= dpcof
        g[i] = (objective(prob, controls, r, target, order, nthreads) - objective(prob, controls, l, target, order, nthreads)) / (2 * dpcof)
    return g


def fill_pq_mat(prob, controls, pcof, t, nderiv):
    """fill_p_mat!/fill_q_mat! at time t -> (p [nderiv,Nc], q [nderiv,Nc]) Taylor-scaled."""
    pk = _pack(prob, controls)
    pc = np.ascontiguousarray(pcof, dtype=np.float64)
    p = np.zeros((nderiv, prob.N_operators), order="F")
    q = np.zeros((nderiv, prob.N_operators), order="F")
    _check(lib().qgdo_fill_pq_mat(pk.ref(), _d(pc), C.c_double(t), C.c_int(nderiv), _d(p), _d(q)))
    return p, q


def eval_pq_derivative(prob, controls, k, pcof_local, t, order, want_grad=True):
    """eval_{p,q}_derivative and eval_grad_{p,q}_derivative! of control k (0-based) on its local slice."""
    pkg = _pkg()
    pk = _pack(prob, controls)
    ctl = pkg.as_control_list(controls)[k]
    pc = np.ascontiguousarray(pcof_local, dtype=np.float64)
    pv, qv = C.c_double(), C.c_double()
    gp = np.zeros(ctl.N_coeff)
    gq = np.zeros(ctl.N_coeff)
    _check(lib().qgdo_eval_pq_derivative(pk.ref(), C.c_int64(k), _d(pc), C.c_double(t), C.c_int(order), C.byref(pv),
                                         C.byref(qv), _d(gp) if want_grad else None, _d(gq) if want_grad else None))
    return pv.value, qv.value, gp, gq


def compute_derivatives(prob, controls, uv, order, cre, cim, adjoint=False):
    pk = _pack(prob, controls)
    uv = np.asfortranarray(uv, dtype=np.float64).copy(order="F")
    cre = np.asfortranarray(cre, dtype=np.float64)
    cim = np.asfortranarray(cim, dtype=np.float64)
    _check(lib().qgdo_compute_derivatives(pk.ref(), _d(uv), C.c_int(order), _d(cre), _d(cim), C.c_int(int(adjoint))))
    return uv


def apply_step_operator(prob, controls, x, order, cre, cim, adjoint=False, lhs=True):
    pk = _pack(prob, controls)
    x = np.ascontiguousarray(x, dtype=np.float64)
    out = np.zeros_like(x)
    cre = np.asfortranarray(cre, dtype=np.float64)
    cim = np.asfortranarray(cim, dtype=np.float64)
    _check(lib().qgdo_apply_step_operator(pk.ref(), _d(x), _d(out), C.c_int(order), _d(cre), _d(cim),
                                          C.c_int(int(adjoint)), C.c_int(int(lhs))))
    return out


def apply_preconditioner(prob, controls, x, order, adjoint=False):
    pk = _pack(prob, controls)
    x = np.ascontiguousarray(x, dtype=np.float64).copy()
    _check(lib().qgdo_apply_preconditioner(pk.ref(), _d(x), C.c_int(order), C.c_int(int(adjoint))))
    return x


def bsplvd(knots, k, x, left, nderiv):
    knots = np.ascontiguousarray(knots, dtype=np.float64)
    out = np.zeros((k, nderiv), order="F")
    _check(lib().qgdo_bsplvd(_d(knots), C.c_int(k), C.c_double(x), C.c_int(left), C.c_int(nderiv), _d(out)))
    return out


def gmres_dense(A, b, x0=None, abstol=0.0, reltol=np.sqrt(np.finfo(float).eps), restart=None, maxiter=None):
    A = np.asfortranarray(A, dtype=np.float64)
    n = A.shape[0]
    b = np.ascontiguousarray(b, dtype=np.float64)
    x = np.zeros(n) if x0 is None else np.ascontiguousarray(x0, dtype=np.float64).copy()
    it = C.c_int()
    _check(lib().qgdo_gmres_dense(_d(A), C.c_int64(n), _d(b), _d(x), C.c_double(abstol), C.c_double(reltol),
                                  C.c_int(restart or min(20, n)), C.c_int(maxiter or n), C.byref(it)))
    return x, it.value
