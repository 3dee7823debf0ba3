"""Latency team (four warps per column) against the one-warp-per-column kernels and the oracle; single-evaluation timing."""
import json
import sys
import time
import numpy as np
sys.path.insert(0, "."); sys.path.insert(0, "oracle")
import __graft_entry__ as g
q = g.load_package()
import oracle as O
def rel(a, b): return float(np.abs(a - b).max() / np.abs(b).max())
res = {}
for name, cfg in (("cnot2", q.configs.cnot2(nsteps=20, tf=20.0, gmres_tol=1e-14)),
                  ("cnot3_333", q.configs.cnot3(nsteps=12, tf=12.0, gmres_tol=1e-14, subsystem_sizes=(3, 3, 3), D1=6)),
                  ("cnot3_444", q.configs.cnot3(nsteps=24, tf=24.0, gmres_tol=1e-14))):
    prob, controls, pcof, target, order = cfg
    tgt = q.complex_to_real(target)
    h = q.Handle(prob, controls)
    h.set_option(q.backend.OPT_LATENCY_TEAM, 2)
    one = h.discrete_adjoint(pcof, tgt, order=order, want_iters=True)
    h.set_option(q.backend.OPT_LATENCY_TEAM, 1)
    team = h.discrete_adjoint(pcof, tgt, order=order, want_iters=True)
    h.close()
    ref = O.discrete_adjoint(prob, controls, pcof, target, order=order)
    res[name] = dict(team_vs_single_grad=rel(team["grad"], one["grad"]), team_vs_oracle_grad=rel(team["grad"][:, 0], ref["grad"]),
                     iters_diff_vs_oracle=int(np.abs(team["iters_fwd"][:, :, 0] - ref["iters_fwd"]).max() + np.abs(team["iters_adj"][:, :, 0] - ref["iters_adj"]).max()),
                     mismatching_solves=int((team["iters_fwd"][:, :, 0] != ref["iters_fwd"]).sum() + (team["iters_adj"][:, :, 0] != ref["iters_adj"]).sum()),
                     infid=abs(team["infidelity"][0] - ref["infidelity"]))
    print(name, res[name], flush=True)
prob, controls, pcof, target, order = q.configs.cnot3(nsteps=550, tf=550.0, gmres_tol=1e-12)
tgt = q.complex_to_real(target)
h = q.Handle(prob, controls)
for mode in (2, 1, 2, 1):
    h.set_option(q.backend.OPT_LATENCY_TEAM, mode)
    ts = []
    for _ in range(3):
        t0 = time.perf_counter(); out = h.discrete_adjoint(pcof, tgt, order=order); ts.append(time.perf_counter() - t0)
    st = h.stats()
    print(json.dumps({"team": mode == 1, "ms_single_eval": min(ts) * 1e3, "k_forward_ms": st["last_forward_ms"], "k_backward_ms": st["last_backward_ms"],
                      "grad0": float(out["grad"][0, 0])}), flush=True)
for B in (4, 9, 18):
    pcs = np.asfortranarray(np.stack([q.configs.cnot3_pcof(len(pcof), s) for s in range(B)], axis=1))
    for mode in (2, 1):
        h.set_option(q.backend.OPT_LATENCY_TEAM, mode)
        h.discrete_adjoint(pcs, tgt, order=order)
        t0 = time.perf_counter(); h.discrete_adjoint(pcs, tgt, order=order); dt = time.perf_counter() - t0
        print(json.dumps({"batch": B, "team": mode == 1, "ms": dt * 1e3}), flush=True)
h.close()
