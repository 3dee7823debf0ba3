#!/bin/bash
# Round 2, run A: parity of the two-pass super-block orthogonalisation, then A/B of the widths at full batch and B = 1.
mkdir -p gpurun_out
V=$PWD/quantumgatedesign.jl_b200/csrc/variants
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02a_pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/r02a_pytest_gpu.log
bash tools/gpu/sweep.sh "- " "QGD_B200_LIB=$V/libqgd_b200_s0.so " "QGD_B200_LIB=$V/libqgd_b200_s64.so " "QGD_B200_LIB=$V/libqgd_b200_s128.so " \
   "- --batch 1" "QGD_B200_LIB=$V/libqgd_b200_s0.so --batch 1" "QGD_B200_LIB=$V/libqgd_b200_s128.so --batch 1" "- --batch 8" "- --batch 74"
for i in 1 2 3 4 5 6 7 8 9; do cp gpurun_out/sweep_$i.json gpurun_out/r02a_sweep_$i.json; done
