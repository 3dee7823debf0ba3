#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_forward -c 1 -o gpurun_out/r02_generic_n125 python tools/gpu/generic_prof.py > gpurun_out/r02m_ncu.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/r02m_ncu.log
