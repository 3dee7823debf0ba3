#!/bin/bash
# Round 2, run G: forced solves on the register-operator sweeps + 4-rows-per-request QR: GPU suite, A/B against the row-at-a-time QR.
mkdir -p gpurun_out
V=$PWD/quantumgatedesign.jl_b200/csrc/variants
timeout 1500 python -m pytest tests -m gpu -x -q --durations=5 > gpurun_out/r02g_pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -12 gpurun_out/r02g_pytest_gpu.log
grep -h "tol 1e-15\|C2 full" gpurun_out/r02g_pytest_gpu.log
bash tools/gpu/sweep.sh "- --no-extras" "QGD_B200_LIB=$V/libqgd_b200_rows1.so --no-extras" "- --no-extras --batch 1" "QGD_B200_LIB=$V/libqgd_b200_rows1.so --no-extras --batch 1" "- --no-extras --batch 296"
for i in 1 2 3 4 5; do cp gpurun_out/sweep_$i.json gpurun_out/r02g_sweep_$i.json; done
