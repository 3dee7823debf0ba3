"""The reference's public functions for the hot path, same names and argument meaning, running on
the B200 through the C ABI.

    eval_forward / eval_forward_      src/forward_evolution.jl:15-70
    discrete_adjoint / discrete_adjoint_   src/eval_grad_discrete_adjoint.jl:83-160
    infidelity_real / infidelity / guard_penalty_real   src/infidelity.jl:7-96

Julia's `f!` (mutating, pre-allocated arrays) is spelled `f_` here.  Arrays are column-major
(`order="F"`) numpy arrays with the reference's shapes.
"""
from __future__ import annotations

import numpy as np

from .backend import get_handle
from .controls import build_control_tables, has_host_controls
from .problem import complex_to_real, real_to_complex


def _check_order(order):
    if order < 2 or order % 2:
        raise ValueError("order must be a positive even integer")


def eval_forward_(uv_history, prob, controls, pcof, order=2, saveEveryNsteps=1, forcing=None, device=-1):
    """eval_forward!(uv_history, prob, controls, pcof; order, saveEveryNsteps, forcing)."""
    _check_order(order)
    m = order // 2
    nsave = prob.nsteps // saveEveryNsteps
    shape = (prob.real_system_size, 1 + m, 1 + nsave, prob.N_initial_conditions)
    assert uv_history.shape == shape, f"uv_history must have shape {shape}"  # @assert of :43
    h = get_handle(prob, controls, device)
    if has_host_controls(controls):  # families the device does not evaluate: host-filled tables (qgd_eval_forward_tables)
        if forcing is not None:
            raise NotImplementedError("forcing with host-evaluated controls")
        cvals, _ = build_control_tables(controls, pcof, prob.tf, prob.nsteps, m)
        out = h.eval_forward_tables(cvals, order=order, save_every=saveEveryNsteps, want_history=True, want_iters=False)
    else:
        out = h.eval_forward(pcof, order=order, save_every=saveEveryNsteps, want_history=True, want_iters=False,
                             forcing=forcing)
    uv_history[...] = out["history"][..., 0]
    return None


def eval_forward(prob, controls, pcof, order=2, saveEveryNsteps=1, forcing=None, device=-1):
    """Return the complex state history [N, 1+nsteps/save, nic] (src/forward_evolution.jl:15-29)."""
    _check_order(order)
    m = order // 2
    nsave = prob.nsteps // saveEveryNsteps
    hist = np.zeros((prob.real_system_size, 1 + m, 1 + nsave, prob.N_initial_conditions), order="F")
    eval_forward_(hist, prob, controls, pcof, order=order, saveEveryNsteps=saveEveryNsteps, forcing=forcing, device=device)
    return real_to_complex(hist[:, 0, :, :])


def discrete_adjoint_(grad, history, lambda_history, adjoint_forcing, prob, controls, pcof, target, order=2,
                      cost_type="Infidelity", history_precomputed=False, device=-1, return_info=False):
    """discrete_adjoint!(grad, history, lambda_history, adjoint_forcing, prob, controls, pcof, target; ...).

    `target` is the COMPLEX N x nic gate (or a real N x nic matrix), passed through complex_to_real as
    in the reference (:126).  history_precomputed=True reuses the history left on the device by the
    previous forward evaluation of the same pcof (the reference passes the array back in)."""
    _check_order(order)
    if cost_type not in ("Infidelity", ":Infidelity"):
        raise ValueError(f"Invalid cost type: {cost_type} (only :Infidelity is on the GPU path, "
                         "src/eval_grad_discrete_adjoint.jl:80)")
    tgt = complex_to_real(target)
    h = get_handle(prob, controls, device)
    if has_host_controls(controls):
        if history is not None or lambda_history is not None or adjoint_forcing is not None or history_precomputed:
            raise NotImplementedError("host-evaluated controls return the gradient, infidelity and guard penalty only")
        cvals, table = build_control_tables(controls, pcof, prob.tf, prob.nsteps, order // 2)
        out = h.discrete_adjoint_tables(cvals, table, tgt, order=order)
        grad[...] = out["grad"][:, 0]
        return (grad, out) if return_info else grad
    out = h.discrete_adjoint(pcof, tgt, order=order, history_precomputed=history_precomputed,
                             want_history=history is not None and not history_precomputed,
                             want_lambda=lambda_history is not None, want_forcing=adjoint_forcing is not None,
                             want_iters=return_info)
    grad[...] = out["grad"][:, 0]
    if history is not None and not history_precomputed:
        history[...] = out["history"][..., 0]
    if lambda_history is not None:
        lambda_history[...] = out["lambda_history"][..., 0]
    if adjoint_forcing is not None:
        adjoint_forcing[...] = out["adjoint_forcing"][..., 0]
    if return_info:
        return grad, out
    return grad


def discrete_adjoint(prob, controls, pcof, target, order=2, cost_type="Infidelity", device=-1):
    """Return the gradient (src/eval_grad_discrete_adjoint.jl:83-102)."""
    grad = np.zeros(len(pcof))
    return discrete_adjoint_(grad, None, None, None, prob, controls, pcof, target, order=order, cost_type=cost_type,
                             device=device)


def eval_grad_forced(prob, controls, pcof, target, order=2, cost_type="Infidelity", device=-1):
    """eval_grad_forced(prob, controls, pcof, target; order, cost_type) (src/eval_grad_forced.jl:18-26): the gradient
    from one forced forward solve per control parameter -- the reference's exactness check of the adjoint."""
    _check_order(order)
    if cost_type not in ("Infidelity", ":Infidelity"):
        raise ValueError(f"Invalid cost type: {cost_type} (only :Infidelity is on the GPU path)")
    h = get_handle(prob, controls, device)
    return h.eval_grad_forced(pcof, complex_to_real(target), order=order)


def discrete_adjoint_batch(prob, controls, pcofs, target, order=2, device=-1, want_iters=False):
    """Many control vectors at once (the batched random-pcof sweep, examples/optimization_with_random_pcof.jl
    style): pcofs [P, B] -> dict(grad [P,B], infidelity [B], guard_penalty [B], ...)."""
    _check_order(order)
    h = get_handle(prob, controls, device)
    return h.discrete_adjoint(np.asarray(pcofs), complex_to_real(target), order=order, want_iters=want_iters)


def infidelity_real(psi, target, N_ess):
    """src/infidelity.jl:7-18 (host arithmetic on a handful of numbers, as in the reference)."""
    psi = np.asarray(psi, dtype=np.float64)
    R = np.asarray(target, dtype=np.float64)
    N = R.shape[0] // 2
    T = np.concatenate([R[N:], -R[:N]], axis=0)
    return 1.0 - (np.sum(psi * R) ** 2 + np.sum(psi * T) ** 2) / N_ess ** 2


def infidelity(prob, controls, pcof, target, order=2, device=-1):
    """src/infidelity.jl:34-47: forward solve on the GPU, then the infidelity of the final state."""
    h = get_handle(prob, controls, device)
    if has_host_controls(controls):
        cvals, _ = build_control_tables(controls, pcof, prob.tf, prob.nsteps, order // 2)
        out = h.eval_forward_tables(cvals, order=order, want_history=False, want_iters=False)
    else:
        out = h.eval_forward(pcof, order=order, want_history=False, want_iters=False)
    return infidelity_real(out["final_state"][:, :, 0], complex_to_real(target), prob.N_ess_levels)


def guard_penalty_real(history, dt, T, W):
    """src/infidelity.jl:56-96 on a host history array [2N, 1+m, 1+nsteps(, nic)]."""
    import scipy.sparse as sp

    h = np.asarray(history)
    if h.ndim == 3:
        h = h[..., None]
    Wd = W.toarray() if sp.issparse(W) else np.asarray(W)
    w0 = h[:, 0, :, :]
    vals = np.einsum("rnc,rs,snc->n", w0, Wd, w0)
    wt = np.ones(w0.shape[1])
    wt[0] = wt[-1] = 0.5
    return float((vals * wt).sum() * dt / T)
