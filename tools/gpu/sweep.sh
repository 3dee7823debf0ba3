#!/bin/bash
# usage: bash tools/gpu/sweep.sh "<env1> <args1>" "<env2> <args2>" ...   each item: ENV=VAL,... then bench args
mkdir -p gpurun_out
i=0
for spec in "$@"; do
  i=$((i+1))
  envs=$(echo "$spec" | cut -d' ' -f1); args=$(echo "$spec" | cut -d' ' -f2-)
  [ "$envs" = "-" ] && envs=""
  env $(echo $envs | tr ',' ' ') timeout 600 python bench.py --steps 2 --warmup 1 --no-cpu-baseline $args > gpurun_out/sweep_$i.json 2> gpurun_out/sweep_$i.err
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/sweep_$i.json").read().strip().splitlines()[-1])
    print("$spec ->", round(d["value"],1), "evals/s  kernel_ms", {k: round(v,1) for k,v in d.get("kernel_ms",{}).items()})
except Exception as e:
    print("$spec -> failed", e)
PY
done
