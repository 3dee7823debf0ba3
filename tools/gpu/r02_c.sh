#!/bin/bash
# Round 2, run C: GPU suite, default bench line (with extras), reference arm.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q --durations=5 > gpurun_out/r02c_pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -12 gpurun_out/r02c_pytest_gpu.log
grep -h "tol 1e-15\|C2 full" gpurun_out/r02c_pytest_gpu.log
( time timeout 900 python bench.py > gpurun_out/r02c_bench.json 2> gpurun_out/r02c_bench.err ) 2>&1 | grep real
tail -3 gpurun_out/r02c_bench.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02c_bench.json").read().strip().splitlines()[-1])
for k in ("value", "ms_per_step", "e2e", "self_check", "latency_ms_single_eval", "latency_detail", "kernel_ms", "gpu_launches"):
    print(k, d.get(k))
print("roofline", {k: d["roofline"][k] for k in ("achieved", "peak", "frac", "kernel_ms")})
print("cpu", d.get("cpu_baseline"))
for k, v in d.get("extra", {}).items():
    print("extra", k, json.dumps(v)[:600])
PY
( time timeout 600 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/r02c_ref.json ) 2>&1 | grep real
cat gpurun_out/r02c_ref.json | cut -c1-600
