"""Tiny row-split problem for compute-sanitizer (racecheck / synccheck / memcheck): (5,4,4) levels, 2 steps, order 4, 2 control vectors."""
import sys
import numpy as np
sys.path.insert(0, ".")
import __graft_entry__ as g
q = g.load_package()
freqs, kerr = q.configs.cnot3_physics()
sizes = (5, 4, 4) if len(sys.argv) < 2 else tuple(int(c) for c in sys.argv[1])
prob = q.DispersiveProblem(sizes, (2, 2, 2), freqs, freqs, kerr, 2.0, 2, sparse_rep=True, gmres_abstol=1e-10, gmres_reltol=1e-10,
                           preconditioner_type=q.DiagonalHamiltonianPreconditioner)
controls = [q.CarrierControl(q.BSpline2Control(4, 2.0), [0.0, -kerr[k, (k + 1) % 3]]) for k in range(3)]
P = q.get_number_of_control_parameters(controls)
pcs = np.asfortranarray(np.stack([q.configs.cnot3_pcof(P, s) for s in range(2)], axis=1))
U0 = q.create_initial_conditions(sizes, (2, 2, 2))
h = q.Handle(prob, controls)
if "--no-tmem" in sys.argv:
    h.set_option(q.backend.OPT_DISABLE_TMEM, 1)
out = h.discrete_adjoint(pcs, q.complex_to_real(U0), order=4)
print("fast launches", h.stats()["fast_path_launches"], "grad norm", float(np.abs(out["grad"]).max()), "finite", bool(np.isfinite(out["grad"]).all()))
h.close()
