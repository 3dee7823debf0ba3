#!/bin/bash
# Round 2, run B: the whole GPU suite (round-2 tests included) + smoke.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q --durations=8 > gpurun_out/r02b_pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -25 gpurun_out/r02b_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5
