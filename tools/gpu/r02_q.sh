#!/bin/bash
# Round 2, run Q (2 GPUs): GPU suite incl. the multi-GPU tests and the bench under torchrun at N = 2 on the final library.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --durations=3 > gpurun_out/r02q_pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -8 gpurun_out/r02q_pytest_gpu.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 2 --warmup 1 > gpurun_out/r02q_bench_2gpu.json 2> gpurun_out/r02q_bench_2gpu.err; echo "torchrun rc=$?"
tail -3 gpurun_out/r02q_bench_2gpu.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02q_bench_2gpu.json").read().strip().splitlines()[-1])
for k in ("value", "ms_per_step", "e2e", "self_check", "nccl_collectives", "columns_sharded", "scaling"):
    print(k, d.get(k))
PY
