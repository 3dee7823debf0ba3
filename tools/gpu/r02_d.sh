#!/bin/bash
# Round 2, run D (2 GPUs): GPU suite incl. the multi-GPU tests, single-evaluation latency with the Hessenberg matrix in
# shared memory, the bench under torchrun at N = 2 (weak line + column-sharded strong-scaling line).
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q --durations=5 > gpurun_out/r02d_pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -12 gpurun_out/r02d_pytest_gpu.log
grep -h "tol 1e-15\|C2 full" gpurun_out/r02d_pytest_gpu.log
bash tools/gpu/sweep.sh "- --batch 1 --no-extras" "- --batch 18 --no-extras" "- --batch 37 --no-extras" "- --no-extras"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 2 --warmup 1 > gpurun_out/r02d_bench_2gpu.json 2> gpurun_out/r02d_bench_2gpu.err; echo "torchrun rc=$?"
tail -3 gpurun_out/r02d_bench_2gpu.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02d_bench_2gpu.json").read().strip().splitlines()[-1])
for k in ("value", "ms_per_step", "e2e", "self_check", "nccl_collectives", "columns_sharded", "scaling", "config"):
    print(k, d.get(k))
PY
NCCL_DEBUG=INFO timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 2 --warmup 1 --shard columns --no-extras > gpurun_out/r02d_bench_2gpu_columns.json 2> gpurun_out/r02d_bench_2gpu_columns.err; echo "torchrun columns rc=$?"
grep -c "NCCL INFO" gpurun_out/r02d_bench_2gpu_columns.err; grep "NCCL INFO.*comm\|NVLS\|Connected all" gpurun_out/r02d_bench_2gpu_columns.err | head -12
tail -c 1500 gpurun_out/r02d_bench_2gpu_columns.json
