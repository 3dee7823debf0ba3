#!/bin/bash
# Round 2, final evidence run on the library with the latency team.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --durations=5 > gpurun_out/r02_final2_pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -10 gpurun_out/r02_final2_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
( time timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_final2_reference_arm.json ) 2>&1 | grep real
( time timeout 1200 python bench.py --steps 3 --warmup 3 > gpurun_out/r02_final2_bench.json 2> gpurun_out/r02_final2_bench.err ) 2>&1 | grep real
tail -2 gpurun_out/r02_final2_bench.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02_final2_bench.json").read().strip().splitlines()[-1])
r = json.loads(open("gpurun_out/r02_final2_reference_arm.json").read().strip().splitlines()[-1])
print("value", d["value"], "e2e", d["e2e"]["value"], "ref", r["value"], "ratio", d["e2e"]["value"] / r["value"], "frac", d["roofline"]["frac"], "lat", d.get("latency_ms_single_eval"), d.get("latency_detail"))
print("same config:", d["config"] == r["config"], "self_check", d["self_check"])
for k, v in d.get("extra", {}).items():
    print("extra", k, json.dumps(v)[:300])
PY
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_forward_fast -c 1 -o gpurun_out/r02_fwd_team_b1 python bench.py --steps 1 --warmup 0 --batch 1 --nsteps 24 --no-extras --no-cpu-baseline > gpurun_out/r02_final2_ncu.log 2>&1; echo "ncu team rc=$?"
timeout 600 python tools/gpu/optimize_time.py | tee gpurun_out/r02_optimize_gate_team.json
