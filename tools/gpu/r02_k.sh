#!/bin/bash
# Round 2, run K: generic sweeps with / without the Krylov prefetch; get_histories with pooled per-level handles.
mkdir -p gpurun_out
V=$PWD/quantumgatedesign.jl_b200/csrc/variants
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02k_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r02k_pytest_gpu.log
echo "prefetch:";  timeout 600 python tools/gpu/generic_time.py | tee gpurun_out/r02k_generic_prefetch.json
echo "no prefetch:"; QGD_B200_LIB=$V/libqgd_b200_nopf.so timeout 600 python tools/gpu/generic_time.py | tee gpurun_out/r02k_generic_noprefetch.json
timeout 600 python - <<'PY' 2>&1 | tee gpurun_out/r02k_get_histories.txt
import time, numpy as np, json
import __graft_entry__ as g
q = g.load_package()
prob, controls, pcof, target, _ = q.configs.cnot3(nsteps=40, tf=40.0, gmres_tol=1e-13)
for conc in (False, True, False, True, False, True):
    t0 = time.perf_counter()
    res = q.get_histories(prob, controls, pcof, 6, orders=(12,), concurrent=conc)
    dt = time.perf_counter() - t0
    s = res["Order 12 (QGD)"]
    print(json.dumps({"concurrent": conc, "seconds": round(dt, 4), "nsteps": s["nsteps"], "device_s_per_level": [round(x, 4) for x in s["elapsed_times"]]}))
PY
