#!/bin/bash
# parity tests, one bench line, and the full-size DRAM traffic of the sweep kernels.  usage: bash tools/gpu/check_dram.sh <tag>
mkdir -p gpurun_out
tag=${1:-chk}
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/${tag}_pytest_gpu.log
bash tools/gpu/sweep.sh "- "
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct --clock-control none -k regex:k_.*ward_fast -c 2 --csv --log-file gpurun_out/${tag}_dram_fullsize.csv python bench.py --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/${tag}_ncu_dram.log 2>&1; echo "dram rc=$?"
grep -E "dram__|duration|hit_rate" gpurun_out/${tag}_dram_fullsize.csv | awk -F'","' '{print substr($5,1,22), $(NF-2), $NF}'
