"""Time the GENERIC row-ELL sweeps (everything that is neither a register-operator nor a dense tensor-core problem):
C2 forced onto them (QGD_OPT_DISABLE_FAST) and a sparse (5,5,5)-level dispersive problem (N = 125; it had no fast path when this was
written -- the row-split groups serve it now, so the problem is forced onto the generic kernels as well; tools/gpu/rs_check.py
compares the two).
usage: python tools/gpu/generic_time.py"""
import json
import sys
import time

import numpy as np

sys.path.insert(0, ".")
import __graft_entry__ as g

q = g.load_package()
res = {}
# C2 on the generic kernels, 74 control vectors x 110 steps
prob, controls, pcof, target, order = q.configs.cnot3(nsteps=110, tf=110.0, gmres_tol=1e-12)
P = len(pcof)
pcs = np.asfortranarray(np.stack([q.configs.cnot3_pcof(P, s) for s in range(74)], axis=1))
h = q.Handle(prob, controls)
h.set_option(q.backend.OPT_DISABLE_FAST, 1)
for rep in range(2):
    out = h.discrete_adjoint(pcs, q.complex_to_real(target), order=order, want_iters=(rep == 1))
st = h.stats()
res["c2_generic"] = dict(batch=74, nsteps=110, fwd_ms=st["last_forward_ms"], bwd_ms=st["last_backward_ms"], fast=st["fast_path_launches"],
                         iters=float(out["iters_fwd"].mean()))
h.set_option(q.backend.OPT_DISABLE_FAST, 0)
for rep in range(2):
    out2 = h.discrete_adjoint(pcs, q.complex_to_real(target), order=order)
st = h.stats()
res["c2_fast"] = dict(fwd_ms=st["last_forward_ms"], bwd_ms=st["last_backward_ms"], fast=st["fast_path_launches"],
                      grad_rel_diff=float(np.abs(out["grad"] - out2["grad"]).max() / np.abs(out2["grad"]).max()))
h.close()
# sparse N = 125: (5,5,5) levels, essential (2,2,2)
freqs, kerr = q.configs.cnot3_physics()
prob = q.DispersiveProblem((5, 5, 5), (2, 2, 2), freqs, freqs, kerr, 60.0, 60, sparse_rep=True, gmres_abstol=1e-12, gmres_reltol=1e-12,
                           preconditioner_type=q.DiagonalHamiltonianPreconditioner)
controls = [q.CarrierControl(q.BSpline2Control(10, 60.0), [0.0, -kerr[k, (k + 1) % 3]]) for k in range(3)]
P = q.get_number_of_control_parameters(controls)
pcs = np.asfortranarray(np.stack([q.configs.cnot3_pcof(P, s) for s in range(74)], axis=1))
U0 = q.create_initial_conditions((5, 5, 5), (2, 2, 2))
h = q.Handle(prob, controls)
h.set_option(q.backend.OPT_DISABLE_FAST, 1)
for rep in range(2):
    out = h.discrete_adjoint(pcs, q.complex_to_real(U0), order=8, want_iters=(rep == 1))
st = h.stats()
res["n125_sparse"] = dict(batch=74, nsteps=60, fwd_ms=st["last_forward_ms"], bwd_ms=st["last_backward_ms"], fast=st["fast_path_launches"],
                          iters=float(out["iters_fwd"].mean()))
h.close()
print(json.dumps(res))
