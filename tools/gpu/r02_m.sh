#!/bin/bash
# Round 2, run M: fast terminal kernel (strict GMRES) -- terminal check, GPU suite, single-evaluation and full-batch timings.
mkdir -p gpurun_out
timeout 300 python tools/gpu/term_check.py 2>&1 | tail -4
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r02m_pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -14 gpurun_out/r02m_pytest_gpu.log
for args in "--batch 1 --no-extras" "--no-extras"; do
  timeout 600 python bench.py --steps 3 --warmup 3 $args 2> /dev/null | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('- $args ->', round(d['value'],1), 'evals/s  kernel_ms', {k: round(v,1) for k,v in d.get('kernel_ms',{}).items()}, 'latency', d.get('latency_ms_single_eval'))
"
done
