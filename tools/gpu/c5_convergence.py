#!/usr/bin/env python
"""C5 (BASELINE.json configs[4]): order-12 Hermite convergence sweep over the time-step count in the style of
get_histories (reference src/Tests/test_convergence.jl:83-93): C2 physics (CNOT3 (4,4,4)/(2,2,2), N = 64, 8 columns) with
smooth degree-16 spline carriers, nsteps = base 2^k with saveEveryNsteps = 2^k; Richardson error and halving rate per
refinement (accuracy), and forward time steps per second for one control vector and for a batch (throughput).
usage: python tools/gpu/c5_convergence.py [tf] [base_nsteps] [levels] [batch]"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from __graft_entry__ import load_package
import numpy as np

q = load_package()
tf = float(sys.argv[1]) if len(sys.argv) > 1 else 40.0
base = int(sys.argv[2]) if len(sys.argv) > 2 else 10
levels = int(sys.argv[3]) if len(sys.argv) > 3 else 7
batch = int(sys.argv[4]) if len(sys.argv) > 4 else 592
order = 12
prob, controls0, _, target, _ = q.configs.cnot3(nsteps=base, tf=tf, gmres_tol=1e-15)
_, kerr = q.configs.cnot3_physics()
controls = []
for k in range(3):
    w = [0.0] + [-kerr[k, l] for l in range(3) if l != k]
    controls.append(q.CarrierControl(q.FortranBSplineControl(16, 20, tf), w))
P = q.get_number_of_control_parameters(controls)
pcof = 0.04 * (np.random.default_rng(0).random(P) - 0.5)

res = dict(config=dict(workload="C5: C2 physics N=64 nic=8 Nc=3, FortranBSpline degree 16 x 3 carriers, order 12", tf=tf, base_nsteps=base,
                       levels=levels, gmres_tol=1e-15), sweep=[])
hist = q.get_histories(prob, controls, pcof, levels, orders=(order,), base_nsteps=base)
s = hist[f"Order {order} (QGD)"]
errs = s["richardson_errors"]
for k in range(len(s["nsteps"])):
    rate = float(np.log2(errs[k - 1] / errs[k])) if k >= 2 and errs[k] > 0 and np.isfinite(errs[k - 1]) else None
    res["sweep"].append(dict(nsteps=int(s["nsteps"][k]), dt=float(s["step_sizes"][k]), richardson_rel_err=None if np.isnan(errs[k]) else float(errs[k]),
                             halving_rate=rate, wall_s=float(s["elapsed_times"][k]),
                             column_steps_per_s=float(s["nsteps"][k] * prob.N_initial_conditions / s["elapsed_times"][k])))
# throughput: a batch of control vectors at the finest useful step size (device time of the forward sweep kernel)
p = prob.copy(); p.nsteps = base * 8; p.gmres_abstol = p.gmres_reltol = 1e-12
h = q.Handle(p, controls)
pcs = np.stack([0.04 * (np.random.default_rng(sd).random(P) - 0.5) for sd in range(batch)], axis=1)
for rep in range(2):
    out = h.eval_forward(pcs, order=order, want_history=False, want_iters=True)
st = h.stats()
res["batched_forward"] = dict(batch=batch, nsteps=p.nsteps, forward_ms=st["last_forward_ms"], fast_path_launches=st["fast_path_launches"],
                              gmres_iters_per_step=float(out["iters"].mean()),
                              column_steps_per_s=float(p.nsteps * 8 * batch / (st["last_forward_ms"] * 1e-3)),
                              forward_solves_per_s=float(batch / (st["last_forward_ms"] * 1e-3)))
h.close()
print(json.dumps(res))
