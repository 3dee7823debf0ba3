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
    ap.add_argument("--no-extras", action="store_true",
                    help="skip the extra measurements of the default single-GPU run (single-evaluation latency, C1, C3, C4, C5)")
    ap.add_argument("--columns-batch", type=int, default=1024,
                    help="control vectors of the column-sharded (strong scaling) measurement: the C3 sweep")
    ap.add_argument("--workload", default="c2", choices=["c2", "c4"],
                    help="c2: the metric's CNOT3 order-8 workload (default); c4: the dense 4-qudit x 4-level shape "
                         "(N=256, 256 columns, order 10, 1000 steps) on the FP64 tensor-core sweeps, --batch control vectors per GPU")
    return ap.parse_args()


def workload(q, nsteps):
    return q.configs.cnot3(nsteps=nsteps, tf=float(nsteps), gmres_tol=1e-12)


def c2_config(nsteps, batch, sharding):
    """`config` of both arms (the reference arm carries the same keys so that the driver's same-config check compares like
    with like; batch_per_gpu there = control vectors per CPU sample)."""
    return {"workload": "C2 CNOT3 (4,4,4)/(2,2,2) N=64 nic=8 Nc=3 order 8 nsteps=%d P=180 gmres_tol=1e-12" % nsteps,
            "batch_per_gpu": batch, "sharding": sharding,
            "l2": "inputs larger than L2 (history + Krylov workspaces of one step >> 126 MB)"}


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
def cpu_sample(q, args, n_parallel, threads_per_eval, first_seed=10_000, want_grads=False):
    """Time `n_parallel` concurrent gradient evaluations of the first `cpu_sample_steps` time steps of the
    C2 workload (same dt, same control functions: prob.tf is shortened, the controls keep tf=nsteps)."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle as O

    prob, controls, pcof, target, order = workload(q, args.nsteps)
    S = min(args.cpu_sample_steps, args.nsteps)
    p = prob.copy()
    p.nsteps = S
    p.tf = prob.tf * S / prob.nsteps
    pcs = pcof_batch(q, len(pcof), n_parallel, first_seed)
    O.lib()
    grads = [None] * n_parallel

    def one(i):
        grads[i] = O.discrete_adjoint(p, controls, pcs[:, i], target, order=order, nthreads=threads_per_eval)["grad"]

    t0 = time.perf_counter()
    ths = [threading.Thread(target=one, args=(i,)) for i in range(n_parallel)]
    for t in ths:
        t.start()
    for t in ths:
        t.join()
    dt = time.perf_counter() - t0
    frac = S / prob.nsteps
    if want_grads:
        return n_parallel * frac / dt, dt, S, grads
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
        "config": c2_config(args.nsteps, args.batch, "none" if args.gpus == 1 else args.shard),
        "note": "CPU restatement of QuantumGateDesign.jl as written (Julia unavailable in this image): the ratio against this arm "
                "depends on the host (%d cores on this box, %d used)" % (cores, n_par * tpe),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": n_par * tpe, "host_cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


# ----------------------------------------------------------------------------------------------------
# GPU arm
# ----------------------------------------------------------------------------------------------------
def time_columns_sharded(q, h, torch, dist, prob, tgt, order, P, Bc, steps, local):
    """Strong scaling of ONE batch of `Bc` control vectors (the C3 sweep) with the initial-condition columns sharded over
    the ranks INSIDE the library: every rank evaluates the same control vectors on its block of columns; per evaluation
    two NCCL all-reduces on the sweep stream (final states between the sweeps, [grad; guard] at the end).
    -> (evals/s, ms per step, collectives per step, max |grad - single-rank grad| / max |grad| for control vector 0)"""
    world = dist.get_world_size()
    q.distributed.attach_library_communicator(h)
    pcs = pcof_batch(q, P, Bc, 0)  # the same control vectors on every rank
    stream = torch.cuda.Stream()
    d_pcof = torch.from_numpy(np.ascontiguousarray(pcs.T)).cuda()
    d_tgt = torch.from_numpy(np.ascontiguousarray(tgt.T)).cuda()
    d_grad = torch.zeros(Bc, P, dtype=torch.float64, device="cuda")
    d_inf = torch.zeros(Bc, dtype=torch.float64, device="cuda")
    d_guard = torch.zeros(Bc, dtype=torch.float64, device="cuda")

    def step():
        h.discrete_adjoint_device(d_pcof.data_ptr(), Bc, d_tgt.data_ptr(), order, d_grad.data_ptr(), d_inf.data_ptr(),
                                  d_guard.data_ptr(), stream.cuda_stream)

    step()
    h.synchronize(stream.cuda_stream)
    coll = h.stats()["collectives"]
    dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(stream):
        e0.record(stream)
        for _ in range(steps):
            step()
        e1.record(stream)
    h.synchronize(stream.cuda_stream)
    dist.barrier()
    t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item()) / steps
    g_sharded = d_grad[0].cpu().numpy()
    h.comm_finalize()
    one = h.discrete_adjoint(pcs[:, 0], tgt, order=order)["grad"][:, 0]   # all columns on this rank
    err = float(np.abs(g_sharded - one).max() / np.abs(one).max())
    return Bc / (ms * 1e-3), ms, int(coll), err


def extras_single_gpu(q, h, args, prob, controls, tgt, target, order, P, local):
    """The other configurations of BASELINE.json, measured by the default single-GPU run so that the driver's record carries
    them (each guarded: a failure here never costs the headline line).  C1: one CNOT2 evaluation; C3: 1024 control vectors
    of C2 in one call; C4: dense N = 256 on the FP64 tensor-core sweeps; C5: order-12 forward batch."""
    ex = {}

    def guarded(name, f):
        try:
            ex[name] = f()
        except Exception as e:  # noqa: BLE001
            ex[name] = {"error": f"{type(e).__name__}: {e}"}

    def c1():
        p1, c1_, pc1, t1, o1 = q.configs.cnot2()
        h1 = q.Handle(p1, c1_, device=local)
        r1 = q.complex_to_real(t1)
        ts = []
        for _ in range(6):
            t0 = time.perf_counter(); h1.discrete_adjoint(pc1, r1, order=o1); ts.append(time.perf_counter() - t0)
        dev = h1.stats()["last_total_ms"]
        h1.close()
        return {"workload": "C1 CNOT2 N=4 nic=4 order 4 nsteps=100 P=40 tol 1e-10, ONE evaluation through the host ABI",
                "ms_per_eval_host_abi": min(ts[1:]) * 1e3, "ms_device": dev}

    def c3():
        pcs = pcof_batch(q, P, 1024, 0)
        h.discrete_adjoint(pcs, tgt, order=order)  # first call at this batch size allocates (23 GB of history)
        t0 = time.perf_counter(); h.discrete_adjoint(pcs, tgt, order=order); dt = time.perf_counter() - t0
        return {"workload": "C3 batched random-pcof sweep: 1024 control vectors of C2 in one call (host ABI)", "seconds": dt,
                "evals_per_s": 1024 / dt}

    def c5():
        p5, c5_, pc5, t5, _ = q.configs.cnot3(nsteps=550, tf=550.0, gmres_tol=1e-12)
        h5 = q.Handle(p5, c5_, device=local)
        pcs = pcof_batch(q, P, 592, 0)
        h5.eval_forward(pcs[:, :8], order=12, want_history=False, want_iters=False)
        h5.eval_forward(pcs, order=12, want_history=False, want_iters=False)
        ms = h5.stats()["last_forward_ms"]
        h5.close()
        return {"workload": "C5 order-12 forward sweep (eval_forward!, as get_histories calls it), C2 physics and controls at dt = 1 "
                            "(about 63 GMRES iterations per step), 592 control vectors x 8 columns x 550 steps", "kernel_ms": ms,
                "column_steps_per_s": 592 * 8 * 550 / (ms * 1e-3)}

    def mid():
        # not a BASELINE configuration: the sparse mid-size family (64 < N <= 256) on the row-split register-operator sweeps,
        # with the generic row-ELL kernels that served it before timed beside them
        freqs, kerr = q.configs.cnot3_physics()
        sizes, nst, nb = (5, 5, 5), 60, 74
        pm = q.DispersiveProblem(sizes, (2, 2, 2), freqs, freqs, kerr, float(nst), nst, sparse_rep=True, gmres_abstol=1e-12, gmres_reltol=1e-12,
                                 preconditioner_type=q.DiagonalHamiltonianPreconditioner)
        cm = [q.CarrierControl(q.BSpline2Control(10, float(nst)), [0.0, -kerr[k, (k + 1) % 3]]) for k in range(3)]
        pcs = pcof_batch(q, q.get_number_of_control_parameters(cm), nb, 0)
        tm = q.complex_to_real(q.create_initial_conditions(sizes, (2, 2, 2)))
        hm = q.Handle(pm, cm, device=local)
        res = {}
        for name, off in (("row_split_groups", 0), ("generic_kernels", 1)):
            hm.set_option(q.backend.OPT_DISABLE_FAST, off)
            for _ in range(2):
                out = hm.discrete_adjoint(pcs, tm, order=8)
            st = hm.stats()
            res[name] = {"forward_ms": st["last_forward_ms"], "adjoint_ms": st["last_backward_ms"], "device_total_ms": st["last_total_ms"],
                         "evals_per_s": nb / (st["last_total_ms"] * 1e-3)}
            res[name + "_grad"] = out["grad"]
        g1, g0 = res.pop("row_split_groups_grad"), res.pop("generic_kernels_grad")
        hm.close()
        return {"workload": "sparse dispersive (5,5,5) levels N=125 nic=8 Nc=3 order 8 nsteps=60 tol 1e-12, 74 control vectors (two warps per column)",
                **res, "grad_rel_diff": float(np.abs(g1 - g0).max() / np.abs(g0).max())}

    guarded("c1", c1)
    guarded("c3", c3)
    guarded("c5", c5)
    guarded("mid_size_n125", mid)
    return ex


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
    pcs = pcof_batch(q, P, B, 0 if shard_columns else rank * B)
    if shard_columns:  # the same control vectors on every rank, columns split inside the library (NCCL on the sweep stream)
        q.distributed.attach_library_communicator(h)

    stream = torch.cuda.Stream()
    d_pcof = torch.from_numpy(np.ascontiguousarray(pcs.T)).cuda()  # [B, P] row-major == [P, B] column-major
    d_tgt = torch.from_numpy(np.ascontiguousarray(tgt.T)).cuda()
    d_grad = torch.zeros(B, P, dtype=torch.float64, device="cuda")
    d_inf = torch.zeros(B, dtype=torch.float64, device="cuda")
    d_guard = torch.zeros(B, dtype=torch.float64, device="cuda")
    launches = collectives = 0

    def step():
        nonlocal launches, collectives
        h.discrete_adjoint_device(d_pcof.data_ptr(), B, d_tgt.data_ptr(), order, d_grad.data_ptr(), d_inf.data_ptr(),
                                  d_guard.data_ptr(), stream.cuda_stream)
        st = h.stats()
        launches += st["kernel_launches"]; collectives += st["collectives"]

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
    h.synchronize(stream.cuda_stream)
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    launches = collectives = 0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    with torch.cuda.stream(stream):
        e0.record(stream)
        for _ in range(args.steps):
            step()
        e1.record(stream)
    h.synchronize(stream.cuda_stream)   # also reports the device error word of the sweeps
    barrier()
    elapsed_ms = e0.elapsed_time(e1)
    clocks = sampler.stop() if rank == 0 else None
    t = torch.tensor([elapsed_ms], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    elapsed_ms = float(t.item())
    total_evals = (B if shard_columns else B * world) * args.steps
    value = total_evals / (elapsed_ms * 1e-3)

    # ---- end-to-end through the host C ABI (host buffers in, gradient / infidelity / guard out), same batch
    e2e_vals = []
    h2d = d2h = 0
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
    e2e_value = (B if shard_columns else B * world) / float(te.item())

    # ---- self-check of the TIMED entry point (outside the timed region): what qgd_discrete_adjoint_device left in d_grad
    # must be bit for bit what the host entry point returns for the same control vectors
    g_dev = d_grad.cpu().numpy().T
    self_check = {"device_entry_equals_host_entry_bitwise": bool(np.array_equal(g_dev, out["grad"]) and
                                                                  np.array_equal(d_inf.cpu().numpy(), out["infidelity"])),
                  "max_abs_diff": float(np.abs(g_dev - out["grad"]).max()), "finite": bool(np.isfinite(g_dev).all())}
    if shard_columns:
        h.comm_finalize()

    # ---- strong scaling of one batch by columns, collectives inside the library (every multi-GPU run reports it)
    columns = None
    if world > 1:
        try:
            v, ms, coll, err = time_columns_sharded(q, h, torch, dist, prob, tgt, order, P, args.columns_batch, max(1, min(args.steps, 2)), local)
            columns = {"sharding": "columns", "scaling": "strong", "batch_total": args.columns_batch, "value": v, "unit": UNIT,
                       "ms_per_step": ms, "nccl_collectives_per_step": coll, "grad_rel_diff_vs_single_rank": err,
                       "exchange": "2 all-reduces per evaluation batch on the sweep stream: final states [2N,nic,B] between the sweeps, "
                                   "[grad; guard] at the end (qgd_comm_init_rank; no host staging)"}
        except Exception as e:  # noqa: BLE001
            columns = {"error": f"{type(e).__name__}: {e}"}

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel, from one instrumented evaluation of the same batch
    rf = {}
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
        traffic = tr.get(f"batch{B}_nsteps{args.nsteps}")
    except Exception:
        pass
    dom = "k_backward" if st["last_backward_ms"] >= st["last_forward_ms"] else "k_forward"
    dom_ms = max(st["last_backward_ms"], st["last_forward_ms"])
    F_dom = work["F_bwd"] if dom == "k_backward" else work["F_fwd"]
    B_dom = work["B_bwd"] if dom == "k_backward" else work["B_fwd"]
    ach_tf = F_dom / (dom_ms * 1e-3) / 1e12
    ach_gb = B_dom / (dom_ms * 1e-3) / 1e9
    traffic_src = ("static: profiles/ncu_traffic.json, the dram__bytes of the committed ncu capture of this same command "
                   "(ncu cannot run inside the timed bench); null when no capture matches this batch / nsteps")
    rf["roofline"] = {"bound": "fp64", "kernel": dom, "achieved": ach_tf, "peak": fp64_peak, "unit": "TFLOP/s",
                      "frac": (ach_tf / fp64_peak) if fp64_peak else None,
                      "traffic": (traffic or {}).get(dom), "traffic_source": traffic_src,
                      "peak_source": "FP64 FMA micro-benchmark run inside this bench (MEASURED_PEAKS.json has no FP64 figure)",
                      "kernel_ms": dom_ms, "algorithmic_flops_per_launch": F_dom}
    rf["roofline_hbm"] = {"bound": "hbm", "kernel": dom, "achieved": ach_gb, "peak": hbm_peak, "unit": "GB/s",
                          "frac": ach_gb / hbm_peak, "traffic": (traffic or {}).get(dom), "peak_source": hbm_src,
                          "algorithmic_bytes_per_launch": B_dom}
    rf["kernel_ms"] = {"k_forward": st["last_forward_ms"], "k_backward": st["last_backward_ms"],
                       "device_total": st["last_total_ms"]}
    rf["gmres_iterations_per_eval"] = {"forward": work["If"] / B, "backward": work["Ib"] / B}
    rf["algorithmic_per_eval"] = {"gflop": work["F_total"] / B / 1e9, "hbm_mbytes": work["B_total"] / B / 1e6}

    # ---- latency of ONE gradient evaluation (what optimize_gate asks for, src/ipopt_optimal_control.jl:257,304): B = 1
    lat = {}
    if world == 1 and not args.no_extras:
        ts = []
        for _ in range(4):
            t0 = time.perf_counter(); h.discrete_adjoint(pcs[:, 0], tgt, order=order); ts.append(time.perf_counter() - t0)
        st1 = h.stats()
        lat = {"latency_ms_single_eval": min(ts[1:]) * 1e3,
               "latency_detail": {"through": "host C ABI, B = 1 (8 columns)", "k_forward_ms": st1["last_forward_ms"],
                                  "k_backward_ms": st1["last_backward_ms"], "device_total_ms": st1["last_total_ms"]}}

    cpu = None
    if not args.no_cpu_baseline and world == 1:
        cores = os.cpu_count() or 1
        tpe = min(prob.N_initial_conditions, cores)
        reps, tot, frac_done = 0, 0.0, 0.0
        parity = None
        while reps < 2 or (tot < 10.0 and reps < 8):  # about 10-30 s of CPU work
            if reps == 0:  # the first sample is control vector 0 of the GPU batch: the oracle also CHECKS the timed path
                v, dt, S, grads = cpu_sample(q, args, 1, tpe, first_seed=0, want_grads=True)
                if S == args.nsteps:
                    parity = float(np.abs(g_dev[:, 0] - grads[0]).max() / np.abs(grads[0]).max())
            else:
                v, dt, S = cpu_sample(q, args, 1, tpe)
            reps += 1; tot += dt; frac_done += S / args.nsteps
        cpu = {"value": frac_done / tot, "unit": UNIT, "cores": tpe, "host_cores": cores, "kind": "port",
               "latency_ms_single_eval": tot / frac_done * 1e3,
               "gpu_gradient_rel_diff_vs_this_port": parity,
               "sample": f"{reps} gradient evaluations of the first {S} of {args.nsteps} time steps of C2, one after the other "
                         f"({tot:.1f} s), {tpe} threads (one per column, like Threads.@threads); C++ restatement of the "
                         "reference as written (Julia unavailable); the first one is control vector 0 of the GPU batch and is "
                         "compared with the gradient the timed entry point produced"}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": elapsed_ms / args.steps, "higher_is_better": True, "scaling": "strong" if shard_columns else "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": c2_config(args.nsteps, B, args.shard if world > 1 else "none"),
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h)},
        "gpu_launches": int(launches), "nccl_collectives": int(collectives),
        "self_check": self_check,
    }
    line.update(rf)
    line.update(lat)
    if columns is not None:
        line["columns_sharded"] = columns
    if cpu:
        line["cpu_baseline"] = cpu
    if world == 1 and not args.no_extras:
        extra = extras_single_gpu(q, h, args, prob, controls, tgt, target, order, P, local)
        h.close()
        try:
            # 9 control vectors = 288 column groups = 1.95 waves of 148 CTAs (4 would leave 20 SMs idle)
            extra["c4"] = c4_measure(q, torch, None, 0, 1, local, nsteps=1000, B=9, steps=1, warmup=1, with_cpu=not args.no_cpu_baseline)
        except Exception as e:  # noqa: BLE001
            extra["c4"] = {"error": f"{type(e).__name__}: {e}"}
        line["extra"] = extra
    print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()


# ----------------------------------------------------------------------------------------------------
# C4: dense Hamiltonian on the FP64 tensor-core sweeps (BASELINE.json configs[3]; not the headline metric's workload)
# ----------------------------------------------------------------------------------------------------
def c4_cpu_sample(q, nsteps_full, ncols=8, nsteps=10):
    """CPU figure beside C4: the oracle's FORWARD sweep (eval_forward!) on `ncols` of the 256 columns for `nsteps` of the
    1000 time steps, scaled by columns x steps (independent, equal-cost units) and doubled for the adjoint sweep.  It is an
    UPPER bound of the CPU evals/s: the reference's adjoint sweep costs 57 instead of 15 operator applications per operator
    evaluation at order 10 (exponential recursions, src/hermite.jl:225-275) and the un-preconditioned terminal-condition
    solves are left out (a full oracle evaluation of 16 columns x 1 step took 592 s on 8 cores when this was written)."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle as O

    prob, controls, pcof, target, order = q.configs.dense_random(N=256, nic=ncols, Nc=4, nsteps=nsteps_full, order=10, gmres_tol=1e-12,
                                                                 dt_norm=1.0, n_basis=20, degree=8)
    p = prob.copy()
    p.nsteps = nsteps
    p.tf = prob.tf * nsteps / nsteps_full
    cores = os.cpu_count() or 1
    t0 = time.perf_counter()
    O.eval_forward(p, controls, pcof, order=order, nthreads=min(ncols, cores))
    dt = time.perf_counter() - t0
    frac = (ncols / 256.0) * (nsteps / float(nsteps_full))
    return {"value": frac / (2.0 * dt), "unit": UNIT, "cores": min(ncols, cores), "host_cores": cores, "kind": "port",
            "sample": f"UPPER bound: oracle forward sweep on {ncols} of 256 columns x {nsteps} of {nsteps_full} time steps of the C4 shape "
                      f"({dt:.1f} s), scaled by columns x steps and doubled for the adjoint sweep (which really costs 57/15 of the forward "
                      "one per operator evaluation); terminal-condition solves not included"}


def c4_measure(q, torch, dist, rank, world, local, nsteps, B, steps, warmup, with_cpu):
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

    for _ in range(warmup):
        h.discrete_adjoint(pcs, tgt, order=order)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    barrier()
    t0 = time.perf_counter()
    dev_ms, launches, out = 0.0, 0, None
    for k in range(steps):  # through the host C ABI: pcof H2D, gradient / infidelity / guard D2H every step
        out = h.discrete_adjoint(pcs, tgt, order=order, want_iters=(k == steps - 1))
        st = h.stats()
        dev_ms += st["last_total_ms"]; launches += st["kernel_launches"]
    barrier()
    wall = time.perf_counter() - t0
    clocks = sampler.stop() if rank == 0 else None
    itf, ita = out["iters_fwd"], out["iters_adj"]
    evals = (2 * nsteps + 1) * 256 * B + itf.sum() + (3 * nsteps - 1) * 256 * B + ita.sum()
    flops_ops = 8.0 * 256 ** 2 * (m * (m + 1) / 2) * float(evals)  # K_d, S_d pre-combined: 8 N^2 per application and column
    # F_alg also holds the gradient inner products F_ip.  Counted in their minimal form -- per time step and column
    # the m Nc products K_k w, S_k w of ONE fused gradient sweep (8 N^2 each), not the reference's 2 Nc m(m+1) mat-vecs per side
    nctrl = 4
    flops_ip = 8.0 * 256 ** 2 * m * nctrl * float(nsteps) * 256 * B
    flops = flops_ops + flops_ip
    t = torch.tensor([wall, dev_ms * 1e-3], dtype=torch.float64, device="cuda")
    if dist:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    wall, dev_s = float(t[0]), float(t[1])
    st = h.stats()
    h.close()
    if rank != 0:
        return None
    try:
        dmma_peak = q.measure_dmma_peak(local)
        peak_src = "FP64 DMMA.8x8x4 micro-benchmark run inside this bench (qgd_measure_dmma_peak; MEASURED_PEAKS.json has no FP64 figure)"
    except Exception:
        dmma_peak, peak_src = 37.1, "fallback: DMMA issue rate of profiles/r01_microbench.txt"
    ach = flops / 1e12 / ((st["last_forward_ms"] + st["last_backward_ms"]) * 1e-3)
    line = {
        "metric": "discrete-adjoint gradient evals/sec, dense N=256 order-10 Hermite (C4)", "value": world * B * steps / dev_s,
        "unit": UNIT, "n_gpus": world, "steps": steps, "warmup": warmup, "ms_per_step": wall / steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"C4 dense random N=256 nic=256 Nc=4 order 10 nsteps={nsteps} P={len(pcof)} gmres_tol=1e-12",
                   "batch_per_gpu": B, "sharding": "pcof" if world > 1 else "none",
                   "l2": "inputs larger than L2 (per-level operators + history of one step >> 126 MB)"},
        "clocks": clocks,
        "e2e": {"value": world * B * steps / wall, "unit": UNIT, "h2d_bytes_per_step": int(st["h2d_bytes"]),
                "d2h_bytes_per_step": int(st["d2h_bytes"])},
        "gpu_launches": launches,
        "roofline": {"bound": "tensor", "kernel": "k_forward_dense + k_backward_dense (FP64 DMMA)", "achieved": ach, "peak": dmma_peak,
                     "unit": "TFLOP/s", "frac": ach / dmma_peak, "traffic": None, "peak_source": peak_src,
                     "algorithmic_flops_per_step": flops,
                     "algorithmic_flops_breakdown": {"operator_applications": flops_ops, "gradient_contractions_minimal_form": flops_ip}},
        "kernel_ms": {"k_forward_dense": st["last_forward_ms"], "k_backward_dense": st["last_backward_ms"], "device_total": st["last_total_ms"]},
        "gmres_iterations_per_step_and_column": {"forward": float(itf.mean()), "backward": float(ita.mean())},
        "cpu_baseline": c4_cpu_sample(q, nsteps) if (with_cpu and world == 1) else None,
    }
    return line


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
    line = c4_measure(q, torch, dist, rank, world, local, nsteps, B, args.steps, args.warmup, not args.no_cpu_baseline)
    if rank == 0:
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
