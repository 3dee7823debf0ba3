#!/usr/bin/env python
"""Time the C4 shape (N = 256 dense, all 256 columns, Nc = 4, order 10) on the generic kernels for a short horizon and
extrapolate to the 1000-step evaluation of SURVEY section 8d.  usage: python tools/gpu/c4_time.py [nsteps] [nic]"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from __graft_entry__ import load_package
import numpy as np

q = load_package()
nsteps = int(sys.argv[1]) if len(sys.argv) > 1 else 8
nic = int(sys.argv[2]) if len(sys.argv) > 2 else 256
batch = int(sys.argv[3]) if len(sys.argv) > 3 else 1
prob, controls, pcof, target, order = q.configs.dense_random(N=256, nic=nic, Nc=4, nsteps=nsteps, order=10, gmres_tol=1e-12,
                                                             dt_norm=1.0, n_basis=20, degree=8)
h = q.Handle(prob, controls)
tgt = q.complex_to_real(target)
t0 = time.perf_counter()
rng = np.random.default_rng(3)
pcs = np.stack([pcof] + [rng.random(len(pcof)) - 0.5 for _ in range(batch - 1)], axis=1)
out = h.discrete_adjoint(pcs, tgt, order=order, want_iters=True)
wall = time.perf_counter() - t0
st = h.stats()
itf, ita = out["iters_fwd"], out["iters_adj"]
m = order // 2
# dense flops: one A_j apply on one column 8 N^2 (K_j, S_j pre-combined), m(m+1)/2 applies per operator evaluation
evals = (2 * nsteps + 1) * nic * batch + itf.sum() + (3 * nsteps - 1) * nic * batch + ita.sum()
flops = 8.0 * 256 ** 2 * (m * (m + 1) / 2) * evals
res = dict(nsteps=nsteps, nic=nic, batch=batch, infidelity=[float(v) for v in out["infidelity"]], grad_norm=float(np.linalg.norm(out["grad"])), wall_s=wall, total_ms=st["last_total_ms"], forward_ms=st["last_forward_ms"], backward_ms=st["last_backward_ms"],
           gmres_iters_per_step_fwd=float(itf.mean()), gmres_iters_per_step_adj=float(ita.mean()),
           fast_path_launches=st["fast_path_launches"], algorithmic_tflop=flops / 1e12,
           achieved_tflops=flops / 1e12 / ((st["last_forward_ms"] + st["last_backward_ms"]) * 1e-3),
           extrapolated_s_per_eval_1000_steps=(st["last_forward_ms"] + st["last_backward_ms"]) * 1e-3 * 1000.0 / nsteps / batch,
           evals_per_s=batch / (st["last_total_ms"] * 1e-3))
print(json.dumps(res))
