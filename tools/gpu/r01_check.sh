set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r01_pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/r01_pytest_gpu.log
timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/r01_bench_b296.json 2> gpurun_out/r01_bench_b296.err; echo "bench rc=$?"
cat gpurun_out/r01_bench_b296.json
timeout 600 python bench.py --steps 3 --warmup 3 --batch 592 --no-cpu-baseline > gpurun_out/r01_bench_b592.json 2> gpurun_out/r01_bench_b592.err; echo "bench rc=$?"
cat gpurun_out/r01_bench_b592.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r01_launches_fast.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_backward_fast -c 1 -o gpurun_out/r01_bwd_fast python bench.py --steps 1 --warmup 0 --batch 148 --nsteps 40 --no-cpu-baseline > gpurun_out/ncu_bwd.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_forward_fast -c 1 -o gpurun_out/r01_fwd_fast python bench.py --steps 1 --warmup 0 --batch 148 --nsteps 40 --no-cpu-baseline > gpurun_out/ncu_fwd.log 2>&1
ls -la gpurun_out
