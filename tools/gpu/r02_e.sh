#!/bin/bash
# Round 2, run E: A/B of kernel variants (tensor-core-only block reduction; 12 / 16 resident warps per SM with spills).
mkdir -p gpurun_out
V=$PWD/quantumgatedesign.jl_b200/csrc/variants
for v in fb67 w12 w16; do
  QGD_B200_LIB=$V/libqgd_b200_$v.so timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "full_cnot3 or gradient_parity or forward_sweep_parity" > gpurun_out/r02e_parity_$v.log 2>&1; echo "parity $v rc=$?"; tail -2 gpurun_out/r02e_parity_$v.log
done
bash tools/gpu/sweep.sh "- --no-extras" "QGD_B200_LIB=$V/libqgd_b200_fd6.so --no-extras" "QGD_B200_LIB=$V/libqgd_b200_bd7.so --no-extras" "QGD_B200_LIB=$V/libqgd_b200_fb67.so --no-extras" \
  "QGD_B200_LIB=$V/libqgd_b200_w12.so --no-extras --batch 888" "QGD_B200_LIB=$V/libqgd_b200_w12.so --no-extras" "QGD_B200_LIB=$V/libqgd_b200_w16.so --no-extras --batch 1184" "QGD_B200_LIB=$V/libqgd_b200_w16.so --no-extras" \
  "QGD_B200_LIB=$V/libqgd_b200_fb67.so --no-extras --batch 1"
for i in 1 2 3 4 5 6 7 8 9; do cp gpurun_out/sweep_$i.json gpurun_out/r02e_sweep_$i.json; done
