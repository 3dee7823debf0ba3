#!/bin/bash
# A/B of kernel variants on one box: parity tests on the default library, then short bench runs of each variant,
# then the full-size DRAM traffic of the sweep kernels.  usage: bash tools/gpu/ab.sh <tag> [variant names...]
mkdir -p gpurun_out
tag=${1:-ab}; shift
V=quantumgatedesign.jl_b200/csrc/variants
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/${tag}_pytest_gpu.log
specs=("- " "QGD_L2_PERSIST=0 ")
for v in "$@"; do specs+=("QGD_B200_LIB=$PWD/$V/libqgd_b200_$v.so "); done
bash tools/gpu/sweep.sh "${specs[@]}"
cp gpurun_out/sweep_1.json gpurun_out/${tag}_bench_main.json
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct --clock-control none -k regex:k_.*ward_fast -c 2 --csv --log-file gpurun_out/${tag}_dram_fullsize.csv python bench.py --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/${tag}_ncu_dram.log 2>&1; echo "dram rc=$?"
grep -E "dram__|duration|hit_rate" gpurun_out/${tag}_dram_fullsize.csv | awk -F'","' '{print $5, $(NF-2), $NF}' | cut -c1-160
