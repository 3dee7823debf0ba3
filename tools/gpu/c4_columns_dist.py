#!/usr/bin/env python
"""One C4 evaluation (N = 256 dense, 256 columns, Nc = 4, order 10) with the initial-condition columns sharded over the
GPUs of one box (ColumnShardedEvaluator: forward sweep per rank, all-gather of the final states, terminal condition +
adjoint sweep per rank, all-reduce of [grad; guard]) -- strong scaling of a single evaluation on the dense tensor-core
sweeps.  Launch: python -m torch.distributed.run --nnodes=1 --nproc-per-node G --master-addr 127.0.0.1 --master-port P
tools/gpu/c4_columns_dist.py [nsteps] [batch]; rank 0 prints one JSON line (time = max over ranks)."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import torch
import torch.distributed as dist
from __graft_entry__ import load_package

q = load_package()
nsteps = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
batch = int(sys.argv[2]) if len(sys.argv) > 2 else 1
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
prob, controls, pcof, target, order = q.configs.dense_random(N=256, nic=256, Nc=4, nsteps=nsteps, order=10, gmres_tol=1e-12,
                                                             dt_norm=1.0, n_basis=20, degree=8)
rng = np.random.default_rng(3)
pcs = np.stack([pcof] + [rng.random(len(pcof)) - 0.5 for _ in range(batch - 1)], axis=1)
tgt = q.complex_to_real(target)
h = q.Handle(prob, controls, device=local)
if world > 1:
    ev = q.distributed.ColumnShardedEvaluator(h, prob.N_initial_conditions, device=torch.device("cuda", local))
    run = lambda: ev.discrete_adjoint(pcs, tgt, order=order)
else:
    run = lambda: h.discrete_adjoint(pcs, tgt, order=order)
times = []
for rep in range(2):  # the first call allocates
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    out = run()
    torch.cuda.synchronize()
    dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
    times.append(float(dt.item()))
if rank == 0:
    print(json.dumps(dict(n_gpus=world, nsteps=nsteps, batch=batch, seconds_per_call=times[-1], first_call_s=times[0],
                          evals_per_s=batch / times[-1], infidelity=[float(v) for v in np.atleast_1d(out["infidelity"])],
                          grad_norm=float(np.linalg.norm(out["grad"])), stats=h.stats())))
if world > 1:
    dist.destroy_process_group()
