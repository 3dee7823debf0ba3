import sys
import numpy as np
sys.path.insert(0, "."); sys.path.insert(0, "oracle")
import __graft_entry__ as g
q = g.load_package()
import oracle as O
def rel(a, b): return float(np.abs(a - b).max() / np.abs(b).max())
freqs, kerr = q.configs.cnot3_physics()
for sizes, order, nsteps in (((5, 5, 5), 8, 5), ((5, 5, 5), 12, 3), ((6, 5, 5), 6, 4), ((6, 6, 6), 8, 3)):
    prob = q.DispersiveProblem(sizes, (2, 2, 2), freqs, freqs, kerr, float(nsteps), nsteps, sparse_rep=True, gmres_abstol=1e-12, gmres_reltol=1e-12,
                               preconditioner_type=q.DiagonalHamiltonianPreconditioner)
    controls = [q.CarrierControl(q.BSpline2Control(6, float(nsteps)), [0.0, -kerr[k, (k + 1) % 3]]) for k in range(3)]
    P = q.get_number_of_control_parameters(controls)
    pcof = q.configs.cnot3_pcof(P, 1)
    U0 = q.create_initial_conditions(sizes, (2, 2, 2))
    tgt = q.complex_to_real(U0)
    ref = O.discrete_adjoint(prob, controls, pcof, U0, order=order)
    h = q.Handle(prob, controls)
    for name, strict in (("rs terminal", 0), ("generic terminal (strict option)", 1)):
        h.set_option(q.backend.OPT_STRICT_MGS, strict)
        out = h.discrete_adjoint(pcof, tgt, order=order, want_iters=True, want_lambda=True)
        lamN = out["lambda_history"][:, 0, -1, :, 0]
        print(sizes, order, name, "iters_term", out["iters_term"][:, 0], "oracle", ref["iters_term"], "cap", 2 * prob.N_tot_levels,
              "lambda_N rel", f"{rel(lamN, ref['lambda_history'][:, 0, -1, :]):.1e}", "grad rel", f"{rel(out['grad'][:, 0], ref['grad']):.1e}", flush=True)
    h.close()
