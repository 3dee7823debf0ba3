"""optimize_gate on top of the GPU hot path -- the production caller of the reference (src/ipopt_optimal_control.jl:187-471).

The reference drives Ipopt's L-BFGS with two closures, `eval_f` (forward solve -> infidelity + guard penalty + ridge penalty,
:243-286) and `eval_grad_f!` (discrete_adjoint!, reusing the state history of `eval_f` when pcof did not change,
`history_precomputed`, :289-346).  Ipopt is not in this image; the same two closures drive scipy's L-BFGS-B here, and the
reuse happens ON THE DEVICE: `eval_f` is one forward sweep + the guard kernel (qgd_adjoint_phase1 on an unsharded handle:
final states and guard penalty out, history resident), `eval_grad_f` is qgd_discrete_adjoint(history_precomputed) on the
resident history when the optimiser asks for the gradient at the point it just evaluated.  Host orchestration only.
"""
from __future__ import annotations

import time

import numpy as np

from .api import infidelity_real
from .backend import get_handle
from .controls import get_number_of_control_parameters, has_host_controls
from .problem import complex_to_real


def optimize_gate(prob, controls, pcof_init, target, order=4, pcof_L=None, pcof_U=None, maxIter=50, ridge_penalty_strength=1e-2,
                  max_cpu_time=60.0 * 60 * 24, device=-1, ftol=0.0, gtol=1e-10):
    """optimize_gate(schro_prob, controls, pcof_init, target; order, pcof_L, pcof_U, maxIter, ridge_penalty_strength, max_cpu_time).
    Returns a dict: final pcof, objective terms, and the per-iteration history the reference's OptimizationHistory records."""
    from scipy.optimize import minimize

    if has_host_controls(controls):
        raise NotImplementedError("optimize_gate: host-evaluated controls go through discrete_adjoint_tables, one call per point")
    N_coeff = get_number_of_control_parameters(controls)
    pcof0 = np.ascontiguousarray(pcof_init, dtype=np.float64)
    assert pcof0.shape == (N_coeff,)
    tgt = complex_to_real(target)
    h = get_handle(prob, controls, device)
    track = dict(last_pcof=None, last_forward_pcof=None, last_adjoint_pcof=None, objective=np.nan, infidelity=np.nan,
                 guard_penalty=np.nan, ridge_penalty=np.nan, grad=None, n_forward=0, n_adjoint=0, n_history_reused=0)
    hist = dict(iter_count=[], elapsed_time=[], objective=[], infidelity=[], guard_penalty=[], ridge_penalty=[], pcof=[])
    t_start = time.perf_counter()

    def terms(pcof, infid, guard):
        ridge = float(pcof @ pcof) * ridge_penalty_strength / len(pcof)
        track.update(infidelity=float(infid), guard_penalty=float(guard), ridge_penalty=ridge, objective=float(infid) + float(guard) + ridge)

    def eval_f(pcof):  # :243-286
        if track["last_pcof"] is None or not np.array_equal(pcof, track["last_pcof"]):
            final, guard = h.adjoint_phase1(pcof, order)  # forward sweep + guard penalty on the device, history stays resident
            terms(pcof, infidelity_real(final[:, :, 0], tgt, prob.N_ess_levels), guard[0])
            track["last_pcof"] = pcof.copy(); track["last_forward_pcof"] = pcof.copy(); track["n_forward"] += 1
        return track["objective"]

    def eval_grad_f(pcof):  # :289-346
        if track["last_adjoint_pcof"] is None or not np.array_equal(pcof, track["last_adjoint_pcof"]):
            reuse = track["last_forward_pcof"] is not None and np.array_equal(pcof, track["last_forward_pcof"])
            out = h.discrete_adjoint(pcof, tgt, order=order, history_precomputed=reuse)
            track["grad"] = out["grad"][:, 0] + 2.0 * ridge_penalty_strength * pcof / len(pcof)
            terms(pcof, out["infidelity"][0], out["guard_penalty"][0])
            track["last_pcof"] = pcof.copy(); track["last_adjoint_pcof"] = pcof.copy(); track["last_forward_pcof"] = pcof.copy()
            track["n_adjoint"] += 1; track["n_history_reused"] += int(reuse)
        return track["grad"]

    def callback(xk):
        eval_f(np.asarray(xk))
        hist["iter_count"].append(len(hist["iter_count"]) + 1)
        hist["elapsed_time"].append(time.perf_counter() - t_start)
        for k in ("objective", "infidelity", "guard_penalty", "ridge_penalty"):
            hist[k].append(track[k])
        hist["pcof"].append(np.array(xk))
        if time.perf_counter() - t_start > max_cpu_time:
            raise StopIteration

    bounds = None
    if pcof_L is not None or pcof_U is not None:
        lo = np.broadcast_to(-np.inf if pcof_L is None else pcof_L, (N_coeff,))
        hi = np.broadcast_to(np.inf if pcof_U is None else pcof_U, (N_coeff,))
        bounds = list(zip(lo, hi))
    f0 = eval_f(pcof0)
    res = minimize(eval_f, pcof0, jac=eval_grad_f, method="L-BFGS-B", bounds=bounds, callback=callback,
                   options=dict(maxiter=int(maxIter), ftol=ftol, gtol=gtol, maxcor=6))
    eval_f(res.x)
    return dict(final_pcof=res.x, initial_objective=f0, final_objective=track["objective"], final_infidelity=track["infidelity"],
                final_guard_penalty=track["guard_penalty"], final_ridge_penalty=track["ridge_penalty"], iterations=int(res.nit),
                n_forward_solves=track["n_forward"], n_adjoint_solves=track["n_adjoint"], n_history_reused=track["n_history_reused"],
                elapsed_time=time.perf_counter() - t_start, optimization_history=hist, scipy_result=res)
