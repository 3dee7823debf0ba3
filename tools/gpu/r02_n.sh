#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --durations=5 > gpurun_out/r02n_pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -14 gpurun_out/r02n_pytest_gpu.log
grep -h "FAILED\|assert" gpurun_out/r02n_pytest_gpu.log | head -20
