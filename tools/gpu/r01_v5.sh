#!/bin/bash
# Round-1 evidence run of the current library: GPU parity tests, bench line (with cpu_baseline), reference arm,
# ncu launch list, full-size DRAM traffic of the sweep kernels, ncu --set full of both sweep kernels (short run).
mkdir -p gpurun_out
tag=${1:-r01_v5}
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/${tag}_pytest_gpu.log
timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; echo "bench rc=$?"
cut -c1-700 gpurun_out/${tag}_bench.json
timeout 300 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/${tag}_bench_reference.json 2> gpurun_out/${tag}_bench_reference.err; echo "ref rc=$?"
cut -c1-300 gpurun_out/${tag}_bench_reference.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/${tag}_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/${tag}_ncu_launch.log 2>&1; echo "launches rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:k_.*ward_fast -c 2 --csv --log-file gpurun_out/${tag}_dram_fullsize.csv python bench.py --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/${tag}_ncu_dram.log 2>&1; echo "dram rc=$?"
bash tools/gpu/prof_bwd.sh ${tag} 296 30
bash tools/gpu/prof_fwd.sh ${tag} 296 30
ls -la gpurun_out | head -40
