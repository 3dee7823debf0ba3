#!/usr/bin/env python
"""bench.py -- discrete-adjoint gradient evals/sec, CNOT3 order-8 Hermite (BASELINE.json metric).

A "step" is one pass of the hot path (forward sweep + guard forcing + terminal condition + backward
adjoint sweep + gradient accumulation, `discrete_adjoint!`) over one batch of `--batch` synthetic
control vectors for the C2 problem of BASELINE.md (N=64, nic=8, Nc=3, nsteps=550, order 8, P=180,
gmres tol 1e-12).  With N GPUs every rank processes its own `--batch` control vectors (pcof sharding,
weak scaling, no data-path collective) -- or, with `--shard columns`, the ranks split the
initial-condition columns of the same control vectors and exchange the final states / gradients.

    python bench.py --gpus 1 --steps 3 --warmup 3
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...
    python bench.py --impl reference ...       # CPU restatement of the reference on the host cores
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
from __graft_entry__ import load_package  # noqa: E402

METRIC = "discrete-adjoint gradient evals/sec, CNOT3 order-8 Hermite"
UNIT = "evals/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=int(os.environ.get("QGD_BENCH_BATCH", "592")),
                    help="control vectors per GPU per step")
    ap.add_argument("--nsteps", type=int, default=550)
    ap.add_argument("--shard", default="pcof", choices=["pcof", "columns"])
    ap.add_argument("--cpu-sample-steps", type=int, default=int(os.environ.get("QGD_CPU_SAMPLE_STEPS", "550")),
                    help="time steps of the C2 problem the CPU baseline integrates per sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--workload", default="c2", choices=["c2", "c4"],
                    help="c2: the metric's CNOT3 order-8 workload (default); c4: the dense 4-qudit x 4-level shape "
                         "(N=256, 256 columns, order 10, 1000 steps) on the FP64 tensor-core sweeps, --batch control vectors per GPU")
    return ap.parse_args()


def workload(q, nsteps):
    return q.configs.cnot3(nsteps=nsteps, tf=float(nsteps), gmres_tol=1e-12)


def pcof_batch(q, P, B, first_seed):
    return np.asfortranarray(np.stack([q.configs.cnot3_pcof(P, first_seed + s) for s in range(B)], axis=1))


# ----------------------------------------------------------------------------------------------------
# algorithmic work
# ----------------------------------------------------------------------------------------------------
def algorithmic_work(prob, order, iters_fwd, iters_adj):
    """-> dict of algorithmic flops/bytes for the evaluations whose iteration counts are given
    (iters_* : [nsteps, nic, B])."""
    m = order // 2
    N, nic, nsteps = prob.N_tot_levels, prob.N_initial_conditions, prob.nsteps
    B = iters_fwd.shape[2]
    a = m * (m + 1) // 2
    nnz_c = sum(op.nnz for op in prob.sym_operators) + sum(op.nnz for op in prob.asym_operators)
    nnz_d = prob.system_sym.nnz + prob.system_asym.nnz
    F_op = 2.0 * (2 * nnz_c) * a + 2.0 * (2 * nnz_d) * m  # one operator evaluation on one column
    If, Ib = float(iters_fwd.sum()), float(iters_adj.sum())

    def mgs(it):
        it = it.astype(np.float64)
        return float((8.0 * N * it * (it + 1) / 2 + 6.0 * N * it).sum())

    cols = nic * B
    F_fwd = F_op * ((2 * nsteps + 1) * cols + If) + mgs(iters_fwd)
    F_ip = cols * nsteps * 2 * prob.N_operators * a * 2 * (2 * (nnz_c / (2 * prob.N_operators)) * 2 + 4 * N)
    F_bwd = F_op * ((3 * nsteps - 1) * cols + Ib) + mgs(iters_adj) + F_ip
    bytes_hist = 8.0 * 2 * N * cols * (nsteps + 1) * (m + 1)
    return dict(F_fwd=F_fwd, F_bwd=F_bwd, F_total=F_fwd + F_bwd, B_fwd=bytes_hist, B_bwd=bytes_hist + 8.0 * 2 * N * cols * (nsteps + 1),
                B_total=8.0 * 2 * N * cols * (nsteps + 1) * (2 * (m + 1) + 1), If=If, Ib=Ib, evals=B)


# ----------------------------------------------------------------------------------------------------
# clocks sampling (B200_PROFILING.md recipe)
# ----------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200",
                                          "-i", str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
            except Exception:
                continue
            names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
            for nm, v in zip(names, r[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ----------------------------------------------------------------------------------------------------
# CPU arm: the oracle (C++ restatement of the reference, as written) on the host cores
# ----------------------------------------------------------------------------------------------------
def cpu_sample(q, args, n_parallel, threads_per_eval):
    """Time `n_parallel` concurrent gradient evaluations of the first `cpu_sample_steps` time steps of the
    C2 workload (same dt, same control functions: prob.tf is shortened, the controls keep tf=nsteps)."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle as O

    prob, controls, pcof, target, order = workload(q, args.nsteps)
    S = min(args.cpu_sample_steps, args.nsteps)
    p = prob.copy()
    p.nsteps = S
    p.tf = prob.tf * S / prob.nsteps
    pcs = pcof_batch(q, len(pcof), n_parallel, 10_000)
    O.lib()

    def one(i):
        O.discrete_adjoint(p, controls, pcs[:, i], target, order=order, nthreads=threads_per_eval)

    t0 = time.perf_counter()
    ths = [threading.Thread(target=one, args=(i,)) for i in range(n_parallel)]
    for t in ths:
        t.start()
    for t in ths:
        t.join()
    dt = time.perf_counter() - t0
    frac = S / prob.nsteps
    return n_parallel * frac / dt, dt, S


def run_reference(args):
    q = load_package()
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    nic = 8
    tpe = min(nic, cores)
    n_par = max(1, cores // tpe)
    vals = []
    for i in range(args.warmup + args.steps):
        v, dt, S = cpu_sample(q, args, n_par, tpe)
        if i >= args.warmup:
            vals.append((v, dt))
    value = float(np.mean([v for v, _ in vals]))
    ms = float(np.mean([d for _, d in vals])) * 1e3
    sample = (f"{n_par} concurrent evals x first {S} of {args.nsteps} time steps of C2, {tpe} threads per eval "
              "(one per column like Threads.@threads), scaled by steps fraction")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": "C2 CNOT3 (4,4,4)/(2,2,2) N=64 nic=8 Nc=3 order 8 nsteps=%d P=180 gmres_tol=1e-12" % args.nsteps,
                   "note": "CPU restatement of QuantumGateDesign.jl as written (Julia unavailable in this image)"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": n_par * tpe, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


# ----------------------------------------------------------------------------------------------------
# GPU arm
# ----------------------------------------------------------------------------------------------------
def run_b200(args):
    import torch

    q = load_package()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the B200 path has no CPU fallback")
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist_

        dist = dist_
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    prob, controls, pcof0, target, order = workload(q, args.nsteps)
    P, B = len(pcof0), args.batch
    tgt = q.complex_to_real(target)
    h = q.Handle(prob, controls, device=local)
    shard_columns = args.shard == "columns" and world > 1
    evaluator = None
    if shard_columns:
        # columns of the same control vectors split over the ranks (quantumgatedesign.jl_b200/distributed.py)
        evaluator = q.distributed.ColumnShardedEvaluator(h, prob.N_initial_conditions, device=torch.device("cuda", local))
        pcs = pcof_batch(q, P, B, 0)  # same control vectors on every rank
    else:
        pcs = pcof_batch(q, P, B, rank * B)

    stream = torch.cuda.Stream()
    d_pcof = torch.from_numpy(np.ascontiguousarray(pcs.T)).cuda()  # [B, P] row-major == [P, B] column-major
    d_tgt = torch.from_numpy(np.ascontiguousarray(tgt.T)).cuda()
    d_grad = torch.zeros(B, P, dtype=torch.float64, device="cuda")
    d_inf = torch.zeros(B, dtype=torch.float64, device="cuda")
    d_guard = torch.zeros(B, dtype=torch.float64, device="cuda")
    launches = 0

    def step_device():
        nonlocal launches
        h.discrete_adjoint_device(d_pcof.data_ptr(), B, d_tgt.data_ptr(), order, d_grad.data_ptr(), d_inf.data_ptr(),
                                  d_guard.data_ptr(), stream.cuda_stream)
        launches += h.stats()["kernel_launches"]

    def step_columns():
        # phase 1 on the owned columns, all-gather of the final states, phase 2, all-reduce of [grad; guard]
        nonlocal launches
        out = evaluator.discrete_adjoint(pcs, tgt, order=order)
        launches += 2 * h.stats()["kernel_launches"]
        return out

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    step = step_columns if shard_columns else step_device
    for _ in range(args.warmup):
        step()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    launches = 0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    fwd_ms, bwd_ms = [], []
    barrier()
    t_wall0 = time.perf_counter()
    with torch.cuda.stream(stream):
        e0.record(stream)
        for _ in range(args.steps):
            step()
        e1.record(stream)
    stream.synchronize()
    barrier()
    t_wall = time.perf_counter() - t_wall0
    elapsed_ms = e0.elapsed_time(e1) if not shard_columns else t_wall * 1e3
    clocks = sampler.stop() if rank == 0 else None
    t = torch.tensor([elapsed_ms], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    elapsed_ms = float(t.item())
    total_evals = (B if shard_columns else B * world) * args.steps
    value = total_evals / (elapsed_ms * 1e-3)

    # ---- end-to-end through the host C ABI (pinned-size host buffers in, gradient out), same batch
    e2e_vals = []
    h2d = d2h = 0
    if not shard_columns:
        for i in range(1 + max(1, min(args.steps, 2))):
            barrier()
            t0 = time.perf_counter()
            out = h.discrete_adjoint(pcs, tgt, order=order)
            dt = time.perf_counter() - t0
            st = h.stats()
            h2d, d2h = st["h2d_bytes"], st["d2h_bytes"]
            if i > 0:
                e2e_vals.append(dt)
        te = torch.tensor([max(e2e_vals)], dtype=torch.float64, device="cuda")
        if dist is not None:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        e2e_value = B * world / float(te.item())
    else:
        e2e_value = value  # the column-sharded step already goes through host buffers and collectives
        h2d = pcs.nbytes + tgt.nbytes
        d2h = 8 * (P * B + 2 * B)

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel, from one instrumented evaluation of the same batch
    rf = {}
    if not shard_columns:
        out = h.discrete_adjoint(pcs, tgt, order=order, want_iters=True)
        st = h.stats()
        work = algorithmic_work(prob, order, out["iters_fwd"], out["iters_adj"])
        try:
            fp64_peak = q.measure_fp64_peak(local)
        except Exception:
            fp64_peak = None
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
        hbm_src = "measured (MEASURED_PEAKS.json)" if peaks else "fallback (B200_PROFILING.md)"
        traffic = None
        try:  # dram bytes of the dominant kernel from the committed `ncu --set full` capture of this same workload
            tr = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
            key = f"batch{B}_nsteps{args.nsteps}"
            traffic = tr.get(key)
        except Exception:
            pass
        dom = "k_backward" if st["last_backward_ms"] >= st["last_forward_ms"] else "k_forward"
        dom_ms = max(st["last_backward_ms"], st["last_forward_ms"])
        F_dom = work["F_bwd"] if dom == "k_backward" else work["F_fwd"]
        B_dom = work["B_bwd"] if dom == "k_backward" else work["B_fwd"]
        ach_tf = F_dom / (dom_ms * 1e-3) / 1e12
        ach_gb = B_dom / (dom_ms * 1e-3) / 1e9
        rf["roofline"] = {"bound": "fp64", "kernel": dom, "achieved": ach_tf, "peak": fp64_peak, "unit": "TFLOP/s",
                          "frac": (ach_tf / fp64_peak) if fp64_peak else None,
                          "traffic": (traffic or {}).get(dom),
                          "peak_source": "FP64 FMA micro-benchmark run inside this bench (MEASURED_PEAKS.json has no FP64 figure)",
                          "kernel_ms": dom_ms, "algorithmic_flops_per_launch": F_dom}
        rf["roofline_hbm"] = {"bound": "hbm", "kernel": dom, "achieved": ach_gb, "peak": hbm_peak, "unit": "GB/s",
                              "frac": ach_gb / hbm_peak, "traffic": (traffic or {}).get(dom), "peak_source": hbm_src,
                              "algorithmic_bytes_per_launch": B_dom}
        rf["kernel_ms"] = {"k_forward": st["last_forward_ms"], "k_backward": st["last_backward_ms"],
                           "device_total": st["last_total_ms"]}
        rf["gmres_iterations_per_eval"] = {"forward": work["If"] / B, "backward": work["Ib"] / B}
        rf["algorithmic_per_eval"] = {"gflop": work["F_total"] / B / 1e9, "hbm_mbytes": work["B_total"] / B / 1e6}

    cpu = None
    if not args.no_cpu_baseline and world == 1:
        cores = os.cpu_count() or 1
        tpe = min(prob.N_initial_conditions, cores)
        reps, tot, frac_done = 0, 0.0, 0.0
        while reps < 2 or (tot < 10.0 and reps < 8):  # about 10-30 s of CPU work
            v, dt, S = cpu_sample(q, args, 1, tpe)
            reps += 1; tot += dt; frac_done += S / args.nsteps
        cpu = {"value": frac_done / tot, "unit": UNIT, "cores": tpe, "kind": "port",
               "sample": f"{reps} gradient evaluations of the first {S} of {args.nsteps} time steps of C2, one after the other "
                         f"({tot:.1f} s), {tpe} threads (one per column, like Threads.@threads); C++ restatement of the "
                         "reference as written (Julia unavailable)"}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": elapsed_ms / args.steps, "higher_is_better": True, "scaling": "strong" if shard_columns else "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "C2 CNOT3 (4,4,4)/(2,2,2) N=64 nic=8 Nc=3 order 8 nsteps=%d P=180 gmres_tol=1e-12" % args.nsteps,
                   "batch_per_gpu": B, "sharding": args.shard if world > 1 else "none",
                   "l2": "inputs larger than L2 (history + Krylov workspaces of one step >> 126 MB)"},
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h)},
        "gpu_launches": int(launches),
    }
    line.update(rf)
    if cpu:
        line["cpu_baseline"] = cpu
    print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()



# ----------------------------------------------------------------------------------------------------
# C4: dense Hamiltonian on the FP64 tensor-core sweeps (BASELINE.json configs[3]; not the headline metric's workload)
# ----------------------------------------------------------------------------------------------------
DMMA_PEAK_TFLOPS = 37.1  # measured DMMA.8x8x4 issue rate on B200, 72.5 warp-instr/ns x 512 flop (profiles/r01_microbench.txt)


def run_c4(args):
    import torch

    q = load_package()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the B200 path has no CPU fallback")
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist_

        dist = dist_
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    nsteps = 1000 if args.nsteps == 550 else args.nsteps
    B = 4 if args.batch == 592 else args.batch
    prob, controls, pcof, target, order = q.configs.dense_random(N=256, nic=256, Nc=4, nsteps=nsteps, order=10, gmres_tol=1e-12,
                                                                 dt_norm=1.0, n_basis=20, degree=8)
    m = order // 2
    rng = np.random.default_rng(100 + rank)  # control vectors shard over the ranks: no data-path collective
    pcs = np.asfortranarray(np.stack([pcof if (rank == 0 and b == 0) else rng.random(len(pcof)) - 0.5 for b in range(B)], axis=1))
    tgt = q.complex_to_real(target)
    h = q.Handle(prob, controls, device=local)

    def barrier():
        torch.cuda.synchronize()
        if dist:
            dist.barrier()

    for _ in range(args.warmup):
        h.discrete_adjoint(pcs, tgt, order=order)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    barrier()
    t0 = time.perf_counter()
    dev_ms, launches, iters = 0.0, 0, None
    for k in range(args.steps):  # through the host C ABI: pcof H2D, gradient / infidelity / guard D2H every step
        out = h.discrete_adjoint(pcs, tgt, order=order, want_iters=(k == args.steps - 1))
        st = h.stats()
        dev_ms += st["last_total_ms"]; launches += st["kernel_launches"]
    barrier()
    wall = time.perf_counter() - t0
    clocks = sampler.stop() if rank == 0 else None
    itf, ita = out["iters_fwd"], out["iters_adj"]
    evals = (2 * nsteps + 1) * 256 * B + itf.sum() + (3 * nsteps - 1) * 256 * B + ita.sum()
    flops = 8.0 * 256 ** 2 * (m * (m + 1) / 2) * float(evals)  # K_d, S_d pre-combined: 8 N^2 per application and column
    t = torch.tensor([wall, dev_ms * 1e-3], dtype=torch.float64, device="cuda")
    if dist:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    wall, dev_s = float(t[0]), float(t[1])
    if rank == 0:
        st = h.stats()
        line = {
            "metric": "discrete-adjoint gradient evals/sec, dense N=256 order-10 Hermite (C4)", "value": world * B * args.steps / dev_s,
            "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": wall / args.steps * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"C4 dense random N=256 nic=256 Nc=4 order 10 nsteps={nsteps} P={len(pcof)} gmres_tol=1e-12",
                       "batch_per_gpu": B, "sharding": "pcof" if world > 1 else "none",
                       "l2": "inputs larger than L2 (per-level operators + history of one step >> 126 MB)"},
            "clocks": clocks,
            "e2e": {"value": world * B * args.steps / wall, "unit": UNIT, "h2d_bytes_per_step": int(st["h2d_bytes"]),
                    "d2h_bytes_per_step": int(st["d2h_bytes"])},
            "gpu_launches": launches,
            "roofline": {"bound": "tensor", "kernel": "k_forward_dense + k_backward_dense (FP64 DMMA)",
                         "achieved": flops / 1e12 / ((st["last_forward_ms"] + st["last_backward_ms"]) * 1e-3), "peak": DMMA_PEAK_TFLOPS,
                         "unit": "TFLOP/s", "frac": flops / 1e12 / ((st["last_forward_ms"] + st["last_backward_ms"]) * 1e-3) / DMMA_PEAK_TFLOPS,
                         "traffic": None, "peak_source": "FP64 DMMA.8x8x4 issue-rate micro-benchmark (profiles/r01_microbench.txt); MEASURED_PEAKS.json has no FP64 figure",
                         "algorithmic_flops_per_step": flops},
            "kernel_ms": {"k_forward_dense": st["last_forward_ms"], "k_backward_dense": st["last_backward_ms"], "device_total": st["last_total_ms"]},
            "gmres_iterations_per_step_and_column": {"forward": float(itf.mean()), "backward": float(ita.mean())},
            "cpu_baseline": None,
        }
        print(json.dumps(line))
    if dist:
        dist.destroy_process_group()


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    elif args.workload == "c4":
        run_c4(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
