#!/bin/bash
# Round 2, run F: ncu evidence -- launch list of the bench command, full captures of the forward sweep at full occupancy
# and for a single evaluation (one warp per SM), full capture of the adjoint sweep.
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 2 --warmup 1 --no-extras --no-cpu-baseline > gpurun_out/r02f_launch_run.log 2>&1; echo "launch list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_forward_fast -c 1 -o gpurun_out/r02_fwd_full python bench.py --steps 1 --warmup 0 --batch 592 --nsteps 24 --no-extras --no-cpu-baseline > gpurun_out/r02f_ncu1.log 2>&1; echo "fwd full rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_backward_fast -c 1 -o gpurun_out/r02_bwd_full python bench.py --steps 1 --warmup 0 --batch 592 --nsteps 24 --no-extras --no-cpu-baseline > gpurun_out/r02f_ncu2.log 2>&1; echo "bwd full rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_forward_fast -c 1 -o gpurun_out/r02_fwd_b1 python bench.py --steps 1 --warmup 0 --batch 1 --nsteps 24 --no-extras --no-cpu-baseline > gpurun_out/r02f_ncu3.log 2>&1; echo "fwd b1 rc=$?"
ls -la gpurun_out/*.ncu-rep
