"""Step-size studies built on the GPU forward solve (callers of the hot path).

    get_histories                   src/Tests/test_convergence.jl:20-146
    richardson_extrap_sol/_rel_err  src/Tests/test_convergence.jl:238-250
    estimate_N_timesteps            src/calculate_timestep.jl (frequency estimate of the drift + max-amplitude controls)
    estimate_timesteps_per_period   src/calculate_timestep.jl:58-98

Host orchestration only: every forward solve is `eval_forward` on the device; the JLD2 logging of the reference is replaced
by the returned dict.  The refinement levels of one order are independent solves of 8 warps each, so they are ENQUEUED TOGETHER
-- one handle (= one stream) per level through `qgd_eval_forward_async`, collected afterwards -- and overlap on the GPU: the
sweep of an order costs the time of its finest level instead of the sum over the levels (`concurrent=False` restores the
reference's one-after-the-other loop; host-evaluated controls always take it).
"""
from __future__ import annotations

import time
from collections import OrderedDict

import numpy as np

from .api import eval_forward
from .backend import Handle, problem_key
from .controls import GRAPEControl, has_host_controls
from .problem import real_to_complex


def richardson_extrap_sol(Ah, A2h, order):
    """Order `order+1` solution from solutions with step h and 2h (test_convergence.jl:247-250)."""
    n = order
    return ((2.0 ** n) * Ah - A2h) / (2.0 ** n - 1.0)


def richardson_extrap_rel_err(Ah, A2h, order):
    """test_convergence.jl:238-241"""
    sol = richardson_extrap_sol(Ah, A2h, order)
    return float(np.linalg.norm(sol - Ah) / np.linalg.norm(sol))


# One handle (= one stream + its device buffers) per refinement level, kept for the next call on the same problem: creating a
# handle costs ~0.1 s (allocations, preconditioner factors, control table), far more than a level's sweep.
_LEVEL_POOL = {"key": None, "handles": []}


def clear_level_pool():
    for h in _LEVEL_POOL["handles"]:
        h.close()
    _LEVEL_POOL["key"], _LEVEL_POOL["handles"] = None, []


def _forward_levels_concurrent(prob, controls, pcof, order, levels, device):
    """levels: [(nsteps, save_every)] -> [(complex history [N, 1+nsteps/save, nic], device seconds)], all levels in flight at once."""
    key = (problem_key(prob, controls), device)
    if _LEVEL_POOL["key"] != key:
        clear_level_pool()
        _LEVEL_POOL["key"] = key
    pool = _LEVEL_POOL["handles"]
    while len(pool) < len(levels):
        pool.append(Handle(prob, controls, device))
    for h, (nsteps, save) in zip(pool, levels):
        h.set_nsteps(nsteps)
        h.set_gmres_tolerances(prob.gmres_abstol, prob.gmres_reltol)
        h.eval_forward_async(pcof, order=order, save_every=save)
    out = []
    for h, _ in zip(pool, levels):
        r = h.eval_forward_collect()
        out.append((real_to_complex(r["history"][:, 0, :, :, 0]), h.stats()["last_forward_ms"] * 1e-3))
    return out


def get_histories(prob, controls, pcof, N_iterations, orders=(2, 4, 6, 8, 10), min_error_limit=-np.inf,
                  max_error_limit=-np.inf, base_nsteps=None, nsteps_change_factor=2, start_iteration=1, device=-1,
                  concurrent=True):
    """test_convergence.jl:20-146: for every order, N_iterations forward solves with nsteps = base * factor^(k-1) and
    saveEveryNsteps = factor^(k-1) (so all histories share the base time grid), Richardson error between consecutive
    refinements, early exit on precision reached / numerical saturation.  Returns the reference's dict of dicts."""
    p = prob.copy()
    base = prob.nsteps if base_nsteps is None else int(base_nsteps)
    ret = OrderedDict()
    concurrent = concurrent and not has_host_controls(controls)
    for order in orders:
        summary = dict(order=order, nsteps=[], step_sizes=[], elapsed_times=[], histories=[], richardson_errors=[])
        ret[f"Order {order} (QGD)"] = summary
        ks = list(range(start_iteration, N_iterations + 1))
        pre = None
        if concurrent:  # every level of this order in flight at once; the stopping rules below only truncate what is returned
            pre = _forward_levels_concurrent(prob, controls, pcof, order,
                                             [(base * nsteps_change_factor ** (k - 1), nsteps_change_factor ** (k - 1)) for k in ks], device)
        for idx, k in enumerate(ks):
            mult = nsteps_change_factor ** (k - 1)
            p.nsteps = base * mult
            if pre is not None:
                history, elapsed = pre[idx]
            else:
                t0 = time.perf_counter()
                history = eval_forward(p, controls, pcof, order=order, saveEveryNsteps=mult, device=device)
                elapsed = time.perf_counter() - t0
            err = float("nan")
            if summary["histories"]:
                err = richardson_extrap_rel_err(history, summary["histories"][-1], order)
            summary["nsteps"].append(p.nsteps)
            summary["step_sizes"].append(p.tf / p.nsteps)
            summary["elapsed_times"].append(elapsed)
            summary["histories"].append(history)
            summary["richardson_errors"].append(err)
            errs = summary["richardson_errors"]
            if errs[-1] < min_error_limit:
                break
            if len(errs) > 2 and errs[-1] < max_error_limit and errs[-1] > errs[-2] > errs[-3]:
                break
    return ret


def _dense(M):
    return np.asarray(M.toarray() if hasattr(M, "toarray") else M, dtype=np.float64)


def get_shortest_period(prob, max_amplitudes):
    """2 pi / max |eig| of H = K_s + i S_s + sum_k a_k (K_k + i S_k) (src/calculate_timestep.jl:17-33)."""
    H = _dense(prob.system_sym) + 1j * _dense(prob.system_asym)
    for k in range(prob.N_operators):
        H = H + max_amplitudes[k] * _dense(prob.sym_operators[k]) + 1j * max_amplitudes[k] * _dense(prob.asym_operators[k])
    return 2.0 * np.pi / np.abs(np.linalg.eigvals(H)).max()


def estimate_N_timesteps(prob, max_amplitudes, timesteps_per_period=40):
    """src/calculate_timestep.jl:35-45"""
    return int(np.ceil(prob.tf / get_shortest_period(prob, max_amplitudes) * timesteps_per_period))


def estimate_timesteps_per_period(prob, max_amplitudes, order, exponents=range(-3, 7), device=-1):
    """calculate_timestep.jl:58-98: constant max-amplitude GRAPE controls, Richardson error of the final state
    between consecutive doublings of the steps per period.  Returns the list of relative errors."""
    p = prob.copy()
    controls = [GRAPEControl(1, p.tf) for _ in range(p.N_operators)]
    pcof = np.repeat(np.asarray(max_amplitudes, dtype=np.float64), 2)
    finals, errs = [], []
    for e in exponents:
        p.nsteps = estimate_N_timesteps(p, max_amplitudes, 2.0 ** e)
        hist = eval_forward(p, controls, pcof, order=order, device=device)
        finals.append(hist[:, -1, :])
        if len(finals) > 1:
            errs.append(richardson_extrap_rel_err(finals[-1], finals[-2], order))
    return errs
