"""eval_grad_forced on the register-operator sweeps vs the generic sweeps vs the adjoint gradient (debug aid)."""
import sys
import numpy as np
sys.path.insert(0, ".")
import __graft_entry__ as g
q = g.load_package()
def rel(a, b): return float(np.abs(a - b).max() / np.abs(b).max())
for nsteps, sizes in ((48, (4, 4, 4)), (12, (4, 4, 4)), (48, (3, 3, 3)), (48, (2, 2, 2))):
    prob, controls, pcof, target, order = q.configs.cnot3(nsteps=nsteps, tf=float(nsteps), gmres_tol=1e-14, subsystem_sizes=sizes, D1=6)
    tgt = q.complex_to_real(target)
    h = q.Handle(prob, controls)
    ga = h.discrete_adjoint(pcof, tgt, order=order)["grad"][:, 0]
    out = {}
    for seg in (0, 1, 5):
        h.set_option(q.backend.OPT_SEG_STEPS, seg)
        out[seg] = h.eval_grad_forced(pcof, tgt, order=order)
    h.set_option(q.backend.OPT_DISABLE_FAST, 1)
    gg = h.eval_grad_forced(pcof, tgt, order=order)
    ga_g = h.discrete_adjoint(pcof, tgt, order=order)["grad"][:, 0]
    h.close()
    print(nsteps, sizes, "fast forced vs adjoint", [f"{rel(out[s], ga):.2e}" for s in out], "generic forced vs adjoint", f"{rel(gg, ga):.2e}",
          "fast forced vs generic forced", f"{rel(out[0], gg):.2e}", "adjoint fast vs generic", f"{rel(ga, ga_g):.2e}")
