#!/bin/bash
# Round 2, run P: ncu capture of the row-split forward sweep (N = 125, two warps per column), bench with the mid-size extra.
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_forward_fast -c 1 -o gpurun_out/r02_rs_fwd python tools/gpu/generic_prof.py > gpurun_out/r02p_ncu.log 2>&1; echo "ncu rc=$?"; tail -2 gpurun_out/r02p_ncu.log
( time timeout 1200 python bench.py --steps 3 --warmup 3 > gpurun_out/r02p_bench.json 2> gpurun_out/r02p_bench.err ) 2>&1 | grep real
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02p_bench.json").read().strip().splitlines()[-1])
print("value", d["value"], "e2e", d["e2e"]["value"], "frac", d["roofline"]["frac"], "lat", d.get("latency_ms_single_eval"))
for k, v in d.get("extra", {}).items():
    print("extra", k, json.dumps(v)[:700])
PY
