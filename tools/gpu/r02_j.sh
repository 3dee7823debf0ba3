#!/bin/bash
# Round 2, run J: concurrent get_histories levels (test + timing against the sequential loop), GPU suite.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02j_pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -6 gpurun_out/r02j_pytest_gpu.log
timeout 600 python - <<'PY' 2>&1 | tee gpurun_out/r02j_get_histories.txt
import time, numpy as np, json
import __graft_entry__ as g
q = g.load_package()
# C5 shape: order-12 convergence sweep on the C2 physics, 6 levels nsteps = 40 * 2^k, saveEveryNsteps = 2^k (get_histories style)
prob, controls, pcof, target, _ = q.configs.cnot3(nsteps=40, tf=40.0, gmres_tol=1e-13)
for conc in (False, True, False, True):
    t0 = time.perf_counter()
    res = q.get_histories(prob, controls, pcof, 6, orders=(12,), concurrent=conc)
    dt = time.perf_counter() - t0
    s = res["Order 12 (QGD)"]
    print(json.dumps({"concurrent": conc, "seconds": round(dt, 3), "nsteps": s["nsteps"], "richardson_errors": [float(f"{e:.3e}") for e in s["richardson_errors"][1:]]}))
PY
