#!/usr/bin/env python
"""Throughput of the dense tensor-core Taylor recursion (qgd_dense.cu) on the C4 shape: N = 256, Nc = 4, order 10,
`ncols` state columns (default 9472 = 8 x 8 x 148: eight CTAs per SM).  usage: python tools/gpu/dense_dmma_time.py [ncols]"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from __graft_entry__ import load_package
import numpy as np

q = load_package()
ncols = int(sys.argv[1]) if len(sys.argv) > 1 else 9472
order = 10
m = order // 2
prob, controls, pcof, target, _ = q.configs.dense_random(N=256, nic=2, Nc=4, nsteps=2, order=order)
rng = np.random.default_rng(0)
cre = rng.standard_normal((m + 1, 4)); cim = rng.standard_normal((m + 1, 4))
uv = np.zeros((512, m + 1, ncols), order="F")
uv[:, 0, :] = rng.standard_normal((512, ncols))
res = {}
for label, env in (("dmma", None), ("generic", "1")):
    h = q.Handle(prob, controls)
    if env:
        h.set_option(q.backend.OPT_DISABLE_DENSE_DMMA, 1)
    nc = ncols if label == "dmma" else min(ncols, 1184)
    best = 1e30
    for rep in range(3):
        h.compute_derivatives(uv[:, :, :nc], order, cre, cim)
        best = min(best, h.stats()["last_forward_ms"])
    flops = 8.0 * 256 ** 2 * (m * (m + 1) / 2) * nc  # pre-combined operators: 8 N^2 per A_d apply and column
    res[label] = dict(ncols=nc, kernel_ms=best, tflops=flops / (best * 1e-3) / 1e12, path=h.stats()["fast_path_launches"])
    h.close()
print(json.dumps(res))
