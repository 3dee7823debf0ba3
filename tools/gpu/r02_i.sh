#!/bin/bash
# Round 2, run I (8 GPUs): the bench as the driver's SCALE step launches it at N = 8 (weak line + column-sharded line), and the
# multi-GPU tests on 4 of the 8 GPUs.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_round2.py -q -k "multi_gpu or communicator" > gpurun_out/r02i_pytest_multi.log 2>&1; echo "pytest multi rc=$?"; tail -3 gpurun_out/r02i_pytest_multi.log
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 3 --warmup 3 > gpurun_out/r02i_bench_8gpu.json 2> gpurun_out/r02i_bench_8gpu.err ) 2>&1 | grep real; echo "torchrun rc=$?"
tail -2 gpurun_out/r02i_bench_8gpu.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02i_bench_8gpu.json").read().strip().splitlines()[-1])
for k in ("value", "ms_per_step", "e2e", "self_check", "columns_sharded", "scaling", "n_gpus"):
    print(k, d.get(k))
PY
( time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 4 --steps 2 --warmup 3 > gpurun_out/r02i_bench_4gpu.json 2> gpurun_out/r02i_bench_4gpu.err ) 2>&1 | grep real
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02i_bench_4gpu.json").read().strip().splitlines()[-1])
for k in ("value", "ms_per_step", "columns_sharded", "n_gpus"):
    print(k, d.get(k))
PY
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29523 bench.py --gpus 8 --steps 1 --warmup 0 --impl reference | cut -c1-400
