"""Long run of the row-split sweeps: (5,5,5) levels, 296 control vectors x 8 columns (two waves of groups), 300 steps, order 8.
Checks: finite, the first and the last control vector equal their single evaluations bit for bit, error word clear."""
import sys, time
import numpy as np
sys.path.insert(0, ".")
import __graft_entry__ as g
q = g.load_package()
freqs, kerr = q.configs.cnot3_physics()
nst, B = 300, 296
prob = q.DispersiveProblem((5, 5, 5), (2, 2, 2), freqs, freqs, kerr, float(nst), nst, sparse_rep=True, gmres_abstol=1e-12, gmres_reltol=1e-12,
                           preconditioner_type=q.DiagonalHamiltonianPreconditioner)
controls = [q.CarrierControl(q.BSpline2Control(10, float(nst)), [0.0, -kerr[k, (k + 1) % 3]]) for k in range(3)]
P = q.get_number_of_control_parameters(controls)
pcs = np.asfortranarray(np.stack([q.configs.cnot3_pcof(P, s) for s in range(B)], axis=1))
tgt = q.complex_to_real(q.create_initial_conditions((5, 5, 5), (2, 2, 2)))
h = q.Handle(prob, controls)
t0 = time.perf_counter(); out = h.discrete_adjoint(pcs, tgt, order=8); dt = time.perf_counter() - t0
st = h.stats()
ok = bool(np.isfinite(out["grad"]).all())
s0 = h.discrete_adjoint(pcs[:, 0], tgt, order=8); s1 = h.discrete_adjoint(pcs[:, B - 1], tgt, order=8)
print({"seconds": round(dt, 2), "fwd_ms": round(st["last_forward_ms"], 1), "bwd_ms": round(st["last_backward_ms"], 1), "fast": st["fast_path_launches"], "finite": ok,
       "first_equal_single": bool(np.array_equal(out["grad"][:, 0], s0["grad"][:, 0])), "last_equal_single": bool(np.array_equal(out["grad"][:, B - 1], s1["grad"][:, 0])),
       "evals_per_s": round(B / (st["last_total_ms"] * 1e-3), 1)})
h.close()
