#!/bin/bash
# ncu full capture of the backward fast kernel on a short C2 run.  usage: bash tools/gpu/prof_bwd.sh <tag> [batch] [nsteps]
mkdir -p gpurun_out
tag=${1:-prof}; batch=${2:-296}; nsteps=${3:-30}
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_backward_ -c 1 -o gpurun_out/${tag}_bwd python bench.py --steps 1 --warmup 0 --batch $batch --nsteps $nsteps --no-cpu-baseline > gpurun_out/${tag}_ncu_bwd.log 2>&1
tail -2 gpurun_out/${tag}_ncu_bwd.log
