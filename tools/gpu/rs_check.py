"""Row-split register-operator sweeps (64 < N <= 256 sparse): parity against the oracle and the generic kernels, then timing.
usage: python tools/gpu/rs_check.py [--time]"""
import json
import sys
import time

import numpy as np

sys.path.insert(0, "."); sys.path.insert(0, "oracle")
import __graft_entry__ as g

q = g.load_package()
import oracle as O


def rel(a, b):
    return float(np.abs(np.asarray(a) - np.asarray(b)).max() / max(np.abs(np.asarray(b)).max(), 1e-300))


def problem(sizes, nsteps, tol=1e-12, precond=None, D1=6):
    freqs, kerr = q.configs.cnot3_physics()
    ess = (2,) * len(sizes)
    prob = q.DispersiveProblem(sizes, ess, freqs, freqs, kerr, float(nsteps), nsteps, sparse_rep=True, gmres_abstol=tol, gmres_reltol=tol,
                               preconditioner_type=precond or q.DiagonalHamiltonianPreconditioner)
    controls = [q.CarrierControl(q.BSpline2Control(D1, float(nsteps)), [0.0, -kerr[k, (k + 1) % 3]]) for k in range(3)]
    P = q.get_number_of_control_parameters(controls)
    return prob, controls, P, q.create_initial_conditions(sizes, ess)


if "--time" not in sys.argv:
    for sizes, order, nsteps in (((5, 5, 5), 8, 5), ((5, 4, 4), 4, 6), ((6, 6, 6), 8, 3), ((6, 5, 5), 6, 4), ((5, 5, 5), 12, 3)):
        prob, controls, P, U0 = problem(sizes, nsteps)
        pcof = q.configs.cnot3_pcof(P, 1)
        tgt = q.complex_to_real(U0)
        t0 = time.perf_counter()
        ref = O.discrete_adjoint(prob, controls, pcof, U0, order=order)
        t_or = time.perf_counter() - t0
        h = q.Handle(prob, controls)
        res = {}
        for name, off in (("rs", 0), ("generic", 1)):
            h.set_option(q.backend.OPT_DISABLE_FAST, off)
            f0 = h.stats()["fast_path_launches"]
            out = h.discrete_adjoint(pcof, tgt, order=order, want_iters=True)
            res[name] = out
            fast = h.stats()["fast_path_launches"] - f0
            mf = int((out["iters_fwd"][..., 0] != ref["iters_fwd"]).sum()); ma = int((out["iters_adj"][..., 0] != ref["iters_adj"]).sum())
            print(sizes, "order", order, name, "fast launches", fast, "grad rel", f"{rel(out['grad'][:, 0], ref['grad']):.2e}",
                  "infid", f"{abs(out['infidelity'][0] - ref['infidelity']):.1e}", "iter mismatches fwd/adj", mf, ma, "of", ref["iters_fwd"].size,
                  "mean its", float(ref["iters_fwd"].mean()), f"oracle {t_or:.1f}s", flush=True)
        print("   rs vs generic grad", f"{rel(res['rs']['grad'], res['generic']['grad']):.2e}")
        h.close()
else:
    for sizes, B in (((5, 5, 5), 74), ((6, 6, 6), 74), ((5, 5, 5), 1), ((6, 6, 6), 1)):
        prob, controls, P, U0 = problem(sizes, 60, D1=10)
        pcs = np.asfortranarray(np.stack([q.configs.cnot3_pcof(P, s) for s in range(B)], axis=1))
        h = q.Handle(prob, controls)
        r = {}
        for name, off in (("rs", 0), ("generic", 1)):
            h.set_option(q.backend.OPT_DISABLE_FAST, off)
            for rep in range(2):
                out = h.discrete_adjoint(pcs, q.complex_to_real(U0), order=8, want_iters=(rep == 1))
            st = h.stats()
            r[name] = dict(fwd_ms=round(st["last_forward_ms"], 1), bwd_ms=round(st["last_backward_ms"], 1), total_ms=round(st["last_total_ms"], 1),
                           fast=st["fast_path_launches"], its=float(out["iters_fwd"].mean()))
            r[name + "_grad"] = out["grad"]
        print(json.dumps(dict(sizes=sizes, batch=B, nsteps=60, rs=r["rs"], generic=r["generic"], grad_rel=rel(r["rs_grad"], r["generic_grad"]))), flush=True)
        h.close()
