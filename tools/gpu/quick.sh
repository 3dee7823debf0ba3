#!/bin/bash
# quick GPU check: parity tests + one bench line.  usage: bash tools/gpu/quick.sh <tag> [bench args...]
mkdir -p gpurun_out
tag=${1:-quick}; shift
timeout 400 python -m pytest tests -m gpu -x -q -s > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?"
grep -E "passed|failed|error|C2 full" gpurun_out/${tag}_pytest.log | tail -8
timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline "$@" > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; echo "bench rc=$?"
tail -3 gpurun_out/${tag}_bench.err
python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/${tag}_bench.json").read().strip().splitlines()[-1])
    print("value", d["value"], "e2e", d["e2e"]["value"], "kernel_ms", d.get("kernel_ms"), "frac", d["roofline"]["frac"], "clocks", d["clocks"])
except Exception as e:
    print("no bench line", e)
PY
