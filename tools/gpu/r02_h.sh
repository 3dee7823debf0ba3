#!/bin/bash
# Round 2, run H: GPU suite + smoke + the default bench line as the driver runs it + reference arm.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q --durations=5 > gpurun_out/r02h_pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -12 gpurun_out/r02h_pytest_gpu.log
grep -h "tol 1e-15\|C2 full" gpurun_out/r02h_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
( time timeout 1200 python bench.py > gpurun_out/r02h_bench.json 2> gpurun_out/r02h_bench.err ) 2>&1 | grep real
tail -3 gpurun_out/r02h_bench.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02h_bench.json").read().strip().splitlines()[-1])
for k in ("value", "ms_per_step", "e2e", "self_check", "latency_ms_single_eval", "kernel_ms", "gpu_launches"):
    print(k, d.get(k))
print("roofline", {k: d["roofline"][k] for k in ("achieved", "peak", "frac", "kernel_ms")})
print("cpu", {k: d["cpu_baseline"][k] for k in ("value", "cores", "host_cores", "latency_ms_single_eval", "gpu_gradient_rel_diff_vs_this_port")})
for k, v in d.get("extra", {}).items():
    print("extra", k, json.dumps(v)[:900])
PY
