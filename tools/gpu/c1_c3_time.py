#!/usr/bin/env python
"""C1 (BASELINE.json configs[0], examples/cnot2_optimization.jl: 2-qubit CNOT, order 4, ONE eval_grad_discrete_adjoint) on the
GPU path and on the CPU restatement, and C3 (1024 random control vectors of C2 in one call on one GPU; the 8-GPU run shards
128 per GPU).  usage: python tools/gpu/c1_c3_time.py [c3_batch]"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from __graft_entry__ import load_package, ROOT
import numpy as np

q = load_package()
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import oracle as O

res = {}
prob, controls, pcof, target, order = q.configs.cnot2()  # nsteps = 100, tf = 100, order 4, P = 40
tgt = q.complex_to_real(target)
h = q.Handle(prob, controls)
walls = []
for rep in range(5):
    t0 = time.perf_counter()
    out = h.discrete_adjoint(pcof, tgt, order=order, want_iters=True)
    walls.append(time.perf_counter() - t0)
st = h.stats()
t0 = time.perf_counter()
ref = O.discrete_adjoint(prob, controls, pcof, target, order=order)
cpu = time.perf_counter() - t0
res["C1"] = dict(workload="cnot2 N=4 nic=4 Nc=2 order 4 nsteps=100 P=40, one gradient evaluation", gpu_call_ms=min(walls) * 1e3,
                 gpu_device_ms=st["last_total_ms"], cpu_restatement_ms=cpu * 1e3,
                 grad_rel_err=float(np.abs(out["grad"][:, 0] - ref["grad"]).max() / np.abs(ref["grad"]).max()),
                 infidelity_rel_err=float(abs(out["infidelity"][0] - ref["infidelity"]) / abs(ref["infidelity"])),
                 gmres_iterations_equal=bool(np.array_equal(out["iters_fwd"][:, :, 0], ref["iters_fwd"]) and
                                             np.array_equal(out["iters_adj"][:, :, 0], ref["iters_adj"])))
h.close()
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
prob, controls, pcof, target, order = q.configs.cnot3()
tgt = q.complex_to_real(target)
pcs = np.asfortranarray(np.stack([q.configs.cnot3_pcof(len(pcof), s) for s in range(B)], axis=1))
h = q.Handle(prob, controls)
for rep in range(2):
    t0 = time.perf_counter()
    out = h.discrete_adjoint(pcs, tgt, order=order)
    wall = time.perf_counter() - t0
st = h.stats()
res["C3_one_gpu"] = dict(workload=f"{B} random control vectors of C2 (CNOT3 order 8, 550 steps) in one call", call_s=wall,
                         device_ms=st["last_total_ms"], evals_per_s=B / wall, infidelity_min=float(out["infidelity"].min()),
                         infidelity_max=float(out["infidelity"].max()), grad_finite=bool(np.isfinite(out["grad"]).all()))
h.close()
print(json.dumps(res))
