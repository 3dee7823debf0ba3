#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r02o_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r02o_pytest_gpu.log
timeout 600 python bench.py --workload c4 --steps 1 --warmup 0 --batch 2 --nsteps 20 --no-cpu-baseline | cut -c1-700
