"""One short forward sweep of the sparse N = 125 problem (for ncu).  Since the row-split groups exist it runs on the register-operator
sweeps (two warps per column: profiles/r02_ncu_rs_fwd.txt); `--generic` forces the row-ELL kernels it was first written for."""
import sys
import numpy as np
sys.path.insert(0, ".")
import __graft_entry__ as g
q = g.load_package()
freqs, kerr = q.configs.cnot3_physics()
prob = q.DispersiveProblem((5, 5, 5), (2, 2, 2), freqs, freqs, kerr, 6.0, 6, sparse_rep=True, gmres_abstol=1e-12, gmres_reltol=1e-12,
                           preconditioner_type=q.DiagonalHamiltonianPreconditioner)
controls = [q.CarrierControl(q.BSpline2Control(10, 60.0), [0.0, -kerr[k, (k + 1) % 3]]) for k in range(3)]
P = q.get_number_of_control_parameters(controls)
pcs = np.asfortranarray(np.stack([q.configs.cnot3_pcof(P, s) for s in range(74)], axis=1))
h = q.Handle(prob, controls)
if "--generic" in sys.argv:
    h.set_option(q.backend.OPT_DISABLE_FAST, 1)
out = h.eval_forward(pcs, order=8, want_history=False, want_iters=True)
print("iters/step", out["iters"].mean(), "fwd ms", h.stats()["last_forward_ms"])
