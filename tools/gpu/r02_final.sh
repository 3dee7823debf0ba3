#!/bin/bash
# Round 2, final evidence run: GPU suite, smoke, the bench pair as the driver runs it, launch list, full-size DRAM traffic.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q --durations=5 > gpurun_out/r02_final_pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -10 gpurun_out/r02_final_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
( time timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_final_reference_arm.json ) 2>&1 | grep real
( time timeout 1200 python bench.py --steps 3 --warmup 3 > gpurun_out/r02_final_bench.json 2> gpurun_out/r02_final_bench.err ) 2>&1 | grep real
tail -2 gpurun_out/r02_final_bench.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02_final_bench.json").read().strip().splitlines()[-1])
r = json.loads(open("gpurun_out/r02_final_reference_arm.json").read().strip().splitlines()[-1])
print("value", d["value"], "e2e", d["e2e"]["value"], "ref", r["value"], "ratio", d["e2e"]["value"] / r["value"], "frac", d["roofline"]["frac"], "lat", d.get("latency_ms_single_eval"))
print("same config:", d["config"] == r["config"])
PY
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_final.csv python bench.py --steps 2 --warmup 1 --no-extras --no-cpu-baseline > gpurun_out/r02_final_launch_run.log 2>&1; echo "launch list rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct --clock-control none -k regex:k_.*ward_fast -c 2 --csv --log-file gpurun_out/r02_dram_fullsize.csv python bench.py --steps 1 --warmup 0 --no-extras --no-cpu-baseline > gpurun_out/r02_final_ncu_dram.log 2>&1; echo "dram rc=$?"
grep -E "dram__|duration|hit_rate" gpurun_out/r02_dram_fullsize.csv | awk -F'","' '{print substr($5,1,22), $(NF-2), $NF}'
