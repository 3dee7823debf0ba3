#!/usr/bin/env python
"""Forward sweep of the C4 shape (N = 256 dense, all 256 columns, Nc = 4, order 10) on the tensor-core sweep
(k_forward_dense) and on the generic kernels.  usage: python tools/gpu/c4_forward_time.py [nsteps] [nic] [batch]"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from __graft_entry__ import load_package
import numpy as np

q = load_package()
nsteps = int(sys.argv[1]) if len(sys.argv) > 1 else 8
nic = int(sys.argv[2]) if len(sys.argv) > 2 else 256
batch = int(sys.argv[3]) if len(sys.argv) > 3 else 1
prob, controls, pcof, target, order = q.configs.dense_random(N=256, nic=nic, Nc=4, nsteps=nsteps, order=10, gmres_tol=1e-12,
                                                             dt_norm=1.0, n_basis=20, degree=8)
m = order // 2
rng = np.random.default_rng(3)
pcs = np.stack([pcof] + [rng.random(len(pcof)) - 0.5 for _ in range(batch - 1)], axis=1)
res = {}
for name, env in (("dense", None), ("generic", "1")):
    if name == "generic" and len(sys.argv) > 4 and sys.argv[4] == "skip-generic":
        continue
    h = q.Handle(prob, controls)
    if env:
        h.set_option(q.backend.OPT_DISABLE_DENSE_SWEEP, 1)
    for rep in range(2):  # first call allocates
        out = h.eval_forward(pcs, order=order, want_history=False)
    st = h.stats()
    it = out["iters"]
    # one operator evaluation on one column: m(m+1)/2 applications of 8 N^2 flops (K_d, S_d pre-combined)
    evals = ((2 * nsteps + 1) * nic * batch + it.sum())
    flops = 8.0 * 256 ** 2 * (m * (m + 1) / 2) * evals
    # lockstep cost actually executed: max iterations over the 8 columns of a group
    res[name] = dict(forward_ms=st["last_forward_ms"], iters_per_step=float(it.mean()), iters_max=int(it.max()),
                     fast_path_launches=st["fast_path_launches"], algorithmic_tflop=flops / 1e12,
                     achieved_tflops=flops / 1e12 / (st["last_forward_ms"] * 1e-3),
                     final_state_norm=float(np.linalg.norm(out["final_state"])))
    res[name + "_final"] = None
    fs = out["final_state"].copy()
    if name == "dense":
        fs_dense, it_dense = fs, it.copy()
    else:
        res["max_abs_diff_final_state"] = float(np.abs(fs - fs_dense).max())
        res["iters_equal_fraction"] = float(np.mean(it == it_dense))
    h.close()
res = {k: v for k, v in res.items() if v is not None}
res["config"] = dict(N=256, nic=nic, Nc=4, order=order, nsteps=nsteps, batch=batch)
print(json.dumps(res))
