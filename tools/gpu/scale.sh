#!/bin/bash
# weak-scaling bench lines on one box: bash tools/gpu/scale.sh "<N list>"   (pcof-sharded, batch 592 per GPU)
mkdir -p gpurun_out
for n in $1; do
  if [ "$n" = "1" ]; then
    timeout 600 python bench.py --gpus 1 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/scale_$n.json 2> gpurun_out/scale_$n.err
  else
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29520+n)) bench.py --gpus $n --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/scale_$n.json 2> gpurun_out/scale_$n.err
  fi
  echo "N=$n rc=$?"; tail -1 gpurun_out/scale_$n.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['n_gpus'], round(d['value'],1), d['unit'], 'ms/step', round(d['ms_per_step'],1), 'e2e', round(d['e2e']['value'],1), d['clocks'])" || tail -3 gpurun_out/scale_$n.err
done
