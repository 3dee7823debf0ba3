#!/bin/bash
# Round 2, run L: generic sweeps with tensor-core dot-product reductions: parity suite on the variant, then timing.
mkdir -p gpurun_out
V=$PWD/quantumgatedesign.jl_b200/csrc/variants
QGD_B200_LIB=$V/libqgd_b200_gdmma.so timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r02l_pytest_gdmma.log 2>&1; echo "pytest (variant) rc=$?"; tail -8 gpurun_out/r02l_pytest_gdmma.log
echo "dmma reductions:"; QGD_B200_LIB=$V/libqgd_b200_gdmma.so timeout 600 python tools/gpu/generic_time.py | tee gpurun_out/r02l_generic_dmma.json
