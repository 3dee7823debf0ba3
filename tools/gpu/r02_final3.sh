#!/bin/bash
# Round 2, final evidence run on the library with the row-split groups and the register-operator terminal kernels.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --durations=5 > gpurun_out/r02_final3_pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -10 gpurun_out/r02_final3_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
( time timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_final3_reference_arm.json ) 2>&1 | grep real
( time timeout 1200 python bench.py --steps 3 --warmup 3 > gpurun_out/r02_final3_bench.json 2> gpurun_out/r02_final3_bench.err ) 2>&1 | grep real
tail -2 gpurun_out/r02_final3_bench.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02_final3_bench.json").read().strip().splitlines()[-1])
r = json.loads(open("gpurun_out/r02_final3_reference_arm.json").read().strip().splitlines()[-1])
print("value", d["value"], "e2e", d["e2e"]["value"], "ref", r["value"], "ratio", d["e2e"]["value"] / r["value"], "frac", d["roofline"]["frac"], "lat", d.get("latency_ms_single_eval"), d.get("latency_detail"))
print("same config:", d["config"] == r["config"], "self_check", d["self_check"], "gpu_launches", d.get("gpu_launches"))
for k, v in d.get("extra", {}).items():
    print("extra", k, json.dumps(v)[:300])
PY
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_final3.csv python bench.py --steps 2 --warmup 1 --no-extras --no-cpu-baseline > gpurun_out/r02_final3_launches.log 2>&1; echo "launch list rc=$?"
timeout 600 python tools/gpu/optimize_time.py | tee gpurun_out/r02_optimize_gate_final3.json
