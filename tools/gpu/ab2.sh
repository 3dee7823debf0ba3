#!/bin/bash
# A/B round 2: parity, per-variant bench, and DRAM traffic of the sweeps for {L2 window on/off} x {evict-first history on/off}
# on a reduced batch (same per-warp footprint).  usage: bash tools/gpu/ab2.sh <tag>
mkdir -p gpurun_out
tag=${1:-ab2}
V=$PWD/quantumgatedesign.jl_b200/csrc/variants
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/${tag}_pytest_gpu.log
bash tools/gpu/sweep.sh "- " "QGD_B200_LIB=$V/libqgd_b200_fwd0.so " "QGD_B200_LIB=$V/libqgd_b200_nocs.so " "QGD_B200_LIB=$V/libqgd_b200_dotchain.so "
cp gpurun_out/sweep_1.json gpurun_out/${tag}_bench_main.json
i=0
for spec in "QGD_L2_PERSIST=1" "QGD_L2_PERSIST=0" "QGD_L2_PERSIST=1 QGD_B200_LIB=$V/libqgd_b200_nocs.so" "QGD_L2_PERSIST=0 QGD_B200_LIB=$V/libqgd_b200_nocs.so"; do
  i=$((i+1))
  env $spec timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct --clock-control none -k regex:k_.*ward_fast -c 2 --csv --log-file gpurun_out/${tag}_dram_$i.csv python bench.py --steps 1 --warmup 0 --batch 148 --nsteps 110 --no-cpu-baseline > gpurun_out/${tag}_ncu_dram_$i.log 2>&1
  echo "== $spec"
  grep -E "dram__|duration|hit_rate" gpurun_out/${tag}_dram_$i.csv | awk -F'","' '{print substr($5,1,22), $(NF-2), $NF}'
done
