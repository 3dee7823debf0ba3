#!/bin/bash
# 2-GPU checks: pcof-sharded (weak) and column-sharded bench under torchrun.  usage: bash tools/gpu/multi.sh <ngpu>
mkdir -p gpurun_out
n=${1:-2}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --steps 2 --warmup 1 --batch 296 > gpurun_out/multi_pcof_$n.json 2> gpurun_out/multi_pcof_$n.err; echo "pcof rc=$?"
tail -1 gpurun_out/multi_pcof_$n.json | cut -c1-400; tail -3 gpurun_out/multi_pcof_$n.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $n --steps 2 --warmup 1 --batch 296 --shard columns > gpurun_out/multi_cols_$n.json 2> gpurun_out/multi_cols_$n.err; echo "cols rc=$?"
tail -1 gpurun_out/multi_cols_$n.json | cut -c1-400; tail -3 gpurun_out/multi_cols_$n.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29513 bench.py --impl reference --gpus $n --steps 1 --warmup 0 > gpurun_out/multi_ref_$n.json 2> gpurun_out/multi_ref_$n.err; echo "ref rc=$?"
tail -1 gpurun_out/multi_ref_$n.json | cut -c1-300
