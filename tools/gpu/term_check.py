import sys
import numpy as np
sys.path.insert(0, "."); sys.path.insert(0, "oracle")
import __graft_entry__ as g
q = g.load_package()
import oracle as O
def rel(a, b): return float(np.abs(a - b).max() / np.abs(b).max())
for sizes, nsteps in (((4, 4, 4), 6), ((3, 3, 3), 6)):
    prob, controls, pcof, target, order = q.configs.cnot3(nsteps=nsteps, tf=float(nsteps), gmres_tol=1e-14, subsystem_sizes=sizes, D1=6)
    tgt = q.complex_to_real(target)
    ref = O.discrete_adjoint(prob, controls, pcof, target, order=order)
    h = q.Handle(prob, controls)
    for strict in (0, 1):
        h.set_option(q.backend.OPT_STRICT_MGS, strict)
        out = h.discrete_adjoint(pcof, tgt, order=order, want_iters=True, want_lambda=True)
        lamN = out["lambda_history"][:, 0, -1, :, 0]
        print(sizes, "strict", strict, "iters_term", out["iters_term"][:, 0], "oracle", ref["iters_term"], "lambda_N rel", rel(lamN, ref["lambda_history"][:, 0, -1, :]),
              "per column", [f"{rel(lamN[:, c], ref['lambda_history'][:, 0, -1, c]):.1e}" for c in range(lamN.shape[1])], "grad rel", rel(out["grad"][:, 0], ref["grad"]))
    h.close()
