"""ctypes binding of csrc/libqgd_b200.so (include/qgd_b200.h).

The CUDA library is the only compute path: if it is missing or cannot create a handle (no sm_100
GPU) every call raises -- there is no CPU fallback and nothing here imports the oracle.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from . import _abi
from ._abi import c_double_p, c_int64_p

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, "csrc")
# QGD_B200_LIB selects another build of the same library (A/B runs of kernel variants during development)
LIB_PATH = os.environ.get("QGD_B200_LIB") or os.path.join(CSRC, "libqgd_b200.so")
_LIB = None

EXPORTS = [
    "qgd_last_error", "qgd_control_n_coeff", "qgd_problem_n_coeff", "qgd_create", "qgd_destroy", "qgd_set_nsteps",
    "qgd_set_gmres_tolerances", "qgd_set_column_shard", "qgd_eval_forward", "qgd_discrete_adjoint",
    "qgd_discrete_adjoint_device", "qgd_adjoint_phase1", "qgd_adjoint_phase2", "qgd_infidelity_real",
    "qgd_eval_controls", "qgd_compute_derivatives", "qgd_get_stats", "qgd_measure_fp64_peak",
    "qgd_eval_forward_forced", "qgd_eval_grad_forced", "qgd_eval_forward_tables", "qgd_discrete_adjoint_tables",
    "qgd_set_option", "qgd_get_option", "qgd_synchronize", "qgd_comm_set_nccl_library", "qgd_comm_get_unique_id",
    "qgd_comm_init_rank", "qgd_comm_finalize", "qgd_init_multi_gpu", "qgd_multi_n_gpus", "qgd_multi_handle",
    "qgd_multi_set_nsteps", "qgd_multi_set_gmres_tolerances", "qgd_multi_discrete_adjoint", "qgd_multi_destroy",
    "qgd_measure_dmma_peak", "qgd_eval_forward_async", "qgd_eval_forward_collect",
]

# option keys of qgd_set_option (include/qgd_b200.h)
OPT_STRICT_MGS, OPT_DISABLE_FAST, OPT_DISABLE_DENSE_SWEEP, OPT_DISABLE_DENSE_DMMA, OPT_DENSE_TERMINAL = 1, 2, 3, 4, 5
OPT_DISABLE_TMEM, OPT_SEG_STEPS, OPT_L2_PERSIST, OPT_LATENCY_WARPS, OPT_TERMINAL_EXCHANGE, OPT_LATENCY_TEAM = 6, 7, 8, 9, 10, 11
SHARD_COLUMNS, SHARD_CONTROL_VECTORS = 0, 1


class QGDError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"qgd_b200 error {code}: {msg}")
        self.code = code


def build(jobs: int = 8, force: bool = False) -> str:
    """Compile the CUDA library in-tree for sm_100a (nvcc cross-compiles without a GPU)."""
    cmd = ["make", "-C", CSRC, f"-j{jobs}"]
    if force:
        cmd.append("-B")
    subprocess.check_call(cmd, stdout=subprocess.DEVNULL)
    return LIB_PATH


def lib():
    global _LIB
    if _LIB is None:
        if not os.path.exists(LIB_PATH):
            raise QGDError(-2, f"{LIB_PATH} is missing: run __graft_entry__.build() (the CUDA extension is the only "
                               "compute path; there is no CPU fallback)")
        L = C.CDLL(LIB_PATH)
        L.qgd_last_error.restype = C.c_char_p
        L.qgd_control_n_coeff.restype = C.c_int64
        L.qgd_problem_n_coeff.restype = C.c_int64
        L.qgd_create.argtypes = [C.POINTER(_abi.qgd_problem_t), C.c_int, C.POINTER(C.c_void_p)]
        L.qgd_destroy.argtypes = [C.c_void_p]
        L.qgd_set_nsteps.argtypes = [C.c_void_p, C.c_int64]
        L.qgd_set_gmres_tolerances.argtypes = [C.c_void_p, C.c_double, C.c_double]
        L.qgd_set_column_shard.argtypes = [C.c_void_p, C.c_int64, C.c_int64]
        L.qgd_eval_forward.argtypes = [C.c_void_p, c_double_p, C.c_int64, C.c_int32, C.c_int64, c_double_p, c_double_p,
                                       c_int64_p]
        L.qgd_discrete_adjoint.argtypes = [C.c_void_p, c_double_p, C.c_int64, c_double_p, C.c_int32, C.c_int32, c_double_p,
                                           c_double_p, c_double_p, c_double_p, c_double_p, c_double_p, c_int64_p,
                                           c_int64_p, c_int64_p]
        L.qgd_discrete_adjoint_device.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_int32, C.c_void_p,
                                                  C.c_void_p, C.c_void_p, C.c_void_p]
        L.qgd_adjoint_phase1.argtypes = [C.c_void_p, c_double_p, C.c_int64, C.c_int32, c_double_p, c_double_p]
        L.qgd_adjoint_phase2.argtypes = [C.c_void_p, c_double_p, c_double_p, c_double_p, c_double_p]
        L.qgd_infidelity_real.argtypes = [C.c_void_p, c_double_p, c_double_p, C.c_int64, c_double_p]
        L.qgd_eval_controls.argtypes = [C.c_void_p, c_double_p, c_double_p, C.c_int64, C.c_int32, c_double_p, c_double_p,
                                        c_double_p, c_double_p]
        L.qgd_compute_derivatives.argtypes = [C.c_void_p, c_double_p, C.c_int64, C.c_int32, c_double_p, c_double_p,
                                              C.c_int32]
        L.qgd_get_stats.argtypes = [C.c_void_p, C.POINTER(_abi.qgd_stats_t)]
        L.qgd_measure_fp64_peak.argtypes = [C.c_int, c_double_p]
        L.qgd_measure_dmma_peak.argtypes = [C.c_int, c_double_p]
        L.qgd_eval_forward_forced.argtypes = [C.c_void_p, c_double_p, C.c_int64, C.c_int32, C.c_int64, c_double_p,
                                              c_double_p, c_double_p, c_int64_p]
        L.qgd_eval_grad_forced.argtypes = [C.c_void_p, c_double_p, c_double_p, C.c_int32, c_double_p]
        L.qgd_eval_forward_tables.argtypes = [C.c_void_p, C.c_int64, C.c_int32, C.c_int64, c_double_p, c_double_p,
                                              c_double_p, c_int64_p]
        L.qgd_discrete_adjoint_tables.argtypes = [C.c_void_p, C.c_int64, C.c_int32, c_double_p, c_double_p, c_double_p,
                                                  c_double_p, c_double_p, c_double_p]
        L.qgd_eval_forward_async.argtypes = [C.c_void_p, c_double_p, C.c_int64, C.c_int32, C.c_int64, C.c_int32]
        L.qgd_eval_forward_collect.argtypes = [C.c_void_p, c_double_p, c_double_p, c_int64_p]
        L.qgd_set_option.argtypes = [C.c_void_p, C.c_int32, C.c_int64]
        L.qgd_get_option.argtypes = [C.c_void_p, C.c_int32, c_int64_p]
        L.qgd_synchronize.argtypes = [C.c_void_p, C.c_void_p]
        L.qgd_comm_set_nccl_library.argtypes = [C.c_char_p]
        L.qgd_comm_get_unique_id.argtypes = [C.c_char_p]
        L.qgd_comm_init_rank.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_char_p]
        L.qgd_comm_finalize.argtypes = [C.c_void_p]
        L.qgd_init_multi_gpu.argtypes = [C.POINTER(_abi.qgd_problem_t), C.c_int32, C.POINTER(C.c_int32), C.POINTER(C.c_void_p)]
        L.qgd_multi_n_gpus.argtypes = [C.c_void_p]
        L.qgd_multi_handle.argtypes = [C.c_void_p, C.c_int32, C.POINTER(C.c_void_p)]
        L.qgd_multi_set_nsteps.argtypes = [C.c_void_p, C.c_int64]
        L.qgd_multi_set_gmres_tolerances.argtypes = [C.c_void_p, C.c_double, C.c_double]
        L.qgd_multi_discrete_adjoint.argtypes = [C.c_void_p, c_double_p, C.c_int64, c_double_p, C.c_int32, C.c_int32,
                                                 c_double_p, c_double_p, c_double_p]
        L.qgd_multi_destroy.argtypes = [C.c_void_p]
        _LIB = L
    return _LIB


def _check(rc):
    if rc != 0:
        raise QGDError(rc, lib().qgd_last_error().decode())


def _dp(a):
    return None if a is None else a.ctypes.data_as(c_double_p)


def _ip(a):
    return None if a is None else a.ctypes.data_as(c_int64_p)


def measure_fp64_peak(device: int = -1) -> float:
    out = np.zeros(1)
    _check(lib().qgd_measure_fp64_peak(device, _dp(out)))
    return float(out[0])


def measure_dmma_peak(device: int = -1) -> float:
    out = np.zeros(1)
    _check(lib().qgd_measure_dmma_peak(device, _dp(out)))
    return float(out[0])


class Handle:
    """Device-resident problem (qgd_handle_t).  One call in flight per handle."""

    def __init__(self, prob, controls, device: int = -1):
        self._pack = _abi.ProblemPack(prob, controls)
        self._h = C.c_void_p()
        _check(lib().qgd_create(self._pack.ref(), int(device), C.byref(self._h)))
        self.N2 = prob.real_system_size
        self.nic = prob.N_initial_conditions
        self.ncol = self.nic
        self.col0 = 0
        self.nsteps = prob.nsteps
        self.P = self._pack.n_coeff
        self.Nc = prob.N_operators

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            lib().qgd_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- mutable knobs of the reference's SchrodingerProb
    def set_nsteps(self, nsteps):
        _check(lib().qgd_set_nsteps(self._h, int(nsteps)))
        self.nsteps = int(nsteps)

    def set_gmres_tolerances(self, abstol, reltol):
        _check(lib().qgd_set_gmres_tolerances(self._h, float(abstol), float(reltol)))

    def set_option(self, key, value):
        """qgd_set_option: behaviour switches (OPT_* above), e.g. h.set_option(OPT_STRICT_MGS, 1)."""
        _check(lib().qgd_set_option(self._h, int(key), int(value)))

    def get_option(self, key):
        out = np.zeros(1, dtype=np.int64)
        _check(lib().qgd_get_option(self._h, int(key), _ip(out)))
        return int(out[0])

    def synchronize(self, stream_ptr=None):
        """Complete (and check) an asynchronous discrete_adjoint_device call."""
        _check(lib().qgd_synchronize(self._h, C.c_void_p(stream_ptr or 0)))

    # -- multi-GPU, one process per GPU (NCCL inside the library)
    def comm_init_rank(self, n_ranks, rank, unique_id: bytes):
        """Attach an NCCL communicator (collective over the ranks) and take this rank's column block."""
        assert len(unique_id) == 128
        _check(lib().qgd_comm_init_rank(self._h, int(n_ranks), int(rank), unique_id))
        c0, c1 = rank * self.nic // n_ranks, (rank + 1) * self.nic // n_ranks
        self.col0, self.ncol = c0, c1 - c0

    def comm_finalize(self):
        _check(lib().qgd_comm_finalize(self._h))
        self.col0, self.ncol = 0, self.nic

    def set_column_shard(self, col_begin, col_count):
        _check(lib().qgd_set_column_shard(self._h, int(col_begin), int(col_count)))
        self.col0, self.ncol = int(col_begin), int(col_count)

    @staticmethod
    def _pcof(pcof, P):
        pc = np.asarray(pcof, dtype=np.float64)
        if pc.ndim == 1:
            pc = pc[:, None]
        if pc.shape[0] != P:
            raise ValueError(f"pcof has {pc.shape[0]} coefficients, the controls need {P}")
        return np.asfortranarray(pc)

    def eval_forward(self, pcof, order=2, save_every=1, want_history=True, want_iters=True, forcing=None):
        """forcing: [2N, m, 1+nsteps, ncol(, B)] as eval_forward!(...; forcing) takes it, or None."""
        pc = self._pcof(pcof, self.P)
        B, m = pc.shape[1], order // 2
        nslots = 1 + self.nsteps // save_every
        hist = np.zeros((self.N2, 1 + m, nslots, self.ncol, B), order="F") if want_history else None
        final = np.zeros((self.N2, self.ncol, B), order="F")
        iters = np.zeros((self.nsteps, self.ncol, B), dtype=np.int64, order="F") if want_iters else None
        if forcing is None:
            _check(lib().qgd_eval_forward(self._h, _dp(pc), B, int(order), int(save_every), _dp(hist), _dp(final),
                                          _ip(iters)))
        else:
            f = np.asarray(forcing, dtype=np.float64)
            if f.ndim == 4:
                f = f[..., None]
            shape = (self.N2, m, self.nsteps + 1, self.ncol, B)
            if f.shape != shape:
                raise ValueError(f"forcing must have shape {shape[:4]} (+ batch), got {f.shape}")
            f = np.asfortranarray(f)
            _check(lib().qgd_eval_forward_forced(self._h, _dp(pc), B, int(order), int(save_every), _dp(f), _dp(hist),
                                                 _dp(final), _ip(iters)))
        return dict(history=hist, final_state=final, iters=iters)

    def eval_forward_async(self, pcof, order=2, save_every=1, want_iters=False):
        """Enqueue a forward solve on this handle's stream and return at once (qgd_eval_forward_async); several handles
        overlap on the GPU.  Finish with `eval_forward_collect`."""
        pc = self._pcof(pcof, self.P)
        self._pending = (pc.shape[1], int(order), int(save_every), bool(want_iters))
        _check(lib().qgd_eval_forward_async(self._h, _dp(pc), pc.shape[1], int(order), int(save_every), int(bool(want_iters))))

    def eval_forward_collect(self, want_history=True):
        B, order, save_every, want_iters = self._pending
        m = order // 2
        nslots = 1 + self.nsteps // save_every
        hist = np.zeros((self.N2, 1 + m, nslots, self.ncol, B), order="F") if want_history else None
        final = np.zeros((self.N2, self.ncol, B), order="F")
        iters = np.zeros((self.nsteps, self.ncol, B), dtype=np.int64, order="F") if want_iters else None
        _check(lib().qgd_eval_forward_collect(self._h, _dp(hist), _dp(final), _ip(iters)))
        return dict(history=hist, final_state=final, iters=iters)

    # -- host-evaluated controls (QGD_CONTROL_HOST_TABLE): tables instead of pcof
    def eval_forward_tables(self, cvals, order=2, save_every=1, want_history=True, want_iters=True):
        cv = np.asfortranarray(cvals, dtype=np.float64)
        m = order // 2
        B = cv.shape[4]
        assert cv.shape == (self.Nc, m + 1, 2, self.nsteps + 1, B), cv.shape
        nslots = 1 + self.nsteps // save_every
        hist = np.zeros((self.N2, 1 + m, nslots, self.ncol, B), order="F") if want_history else None
        final = np.zeros((self.N2, self.ncol, B), order="F")
        iters = np.zeros((self.nsteps, self.ncol, B), dtype=np.int64, order="F") if want_iters else None
        _check(lib().qgd_eval_forward_tables(self._h, B, int(order), int(save_every), _dp(cv), _dp(hist), _dp(final),
                                             _ip(iters)))
        return dict(history=hist, final_state=final, iters=iters)

    def discrete_adjoint_tables(self, cvals, table, target_real, order=2):
        cv = np.asfortranarray(cvals, dtype=np.float64)
        tb = np.asfortranarray(table, dtype=np.float64)
        m = order // 2
        B = cv.shape[4]
        assert cv.shape == (self.Nc, m + 1, 2, self.nsteps + 1, B), cv.shape
        assert tb.shape == (self.P, m + 1, 2, self.nsteps + 1), tb.shape
        tgt = np.asfortranarray(target_real, dtype=np.float64)
        if tgt.shape != (self.N2, self.nic):
            raise ValueError(f"target must be the real-stacked [2N, nic] = {(self.N2, self.nic)} array, got {tgt.shape}")
        grad = np.zeros((self.P, B), order="F")
        infid = np.zeros(B)
        guard = np.zeros(B)
        _check(lib().qgd_discrete_adjoint_tables(self._h, B, int(order), _dp(cv), _dp(tb), _dp(tgt), _dp(grad), _dp(infid),
                                                 _dp(guard)))
        return dict(grad=grad, infidelity=infid, guard_penalty=guard)

    def eval_grad_forced(self, pcof, target_real, order=2):
        """eval_grad_forced (src/eval_grad_forced.jl:18-195): P forced solves batched on the device."""
        pc = np.ascontiguousarray(pcof, dtype=np.float64)
        if pc.shape != (self.P,):
            raise ValueError(f"pcof has shape {pc.shape}, the controls need ({self.P},)")
        tgt = np.asfortranarray(target_real, dtype=np.float64)
        if tgt.shape != (self.N2, self.nic):
            raise ValueError(f"target must be the real-stacked [2N, nic] = {(self.N2, self.nic)} array, got {tgt.shape}")
        grad = np.zeros(self.P)
        _check(lib().qgd_eval_grad_forced(self._h, _dp(pc), _dp(tgt), int(order), _dp(grad)))
        return grad

    def discrete_adjoint(self, pcof, target_real, order=2, history_precomputed=False, want_history=False,
                         want_lambda=False, want_forcing=False, want_iters=False):
        pc = self._pcof(pcof, self.P)
        B, m, Nt = pc.shape[1], order // 2, self.nsteps + 1
        tgt = np.asfortranarray(target_real, dtype=np.float64)
        if tgt.shape != (self.N2, self.nic):
            raise ValueError(f"target must be the real-stacked [2N, nic] = {(self.N2, self.nic)} array, got {tgt.shape}")
        grad = np.zeros((self.P, B), order="F")
        infid = np.zeros(B)
        guard = np.zeros(B)
        hist = np.zeros((self.N2, 1 + m, Nt, self.ncol, B), order="F") if want_history else None
        lam = np.zeros((self.N2, 1 + m, Nt, self.ncol, B), order="F") if want_lambda else None
        forc = np.zeros((self.N2, Nt, self.ncol, B), order="F") if want_forcing else None
        itf = np.zeros((self.nsteps, self.ncol, B), dtype=np.int64, order="F") if want_iters else None
        ita = np.zeros((self.nsteps, self.ncol, B), dtype=np.int64, order="F") if want_iters else None
        itt = np.zeros((self.nic, B), dtype=np.int64, order="F") if want_iters else None
        _check(lib().qgd_discrete_adjoint(self._h, _dp(pc), B, _dp(tgt), int(order), int(bool(history_precomputed)),
                                          _dp(grad), _dp(infid), _dp(guard), _dp(hist), _dp(lam), _dp(forc), _ip(itf),
                                          _ip(ita), _ip(itt)))
        return dict(grad=grad, infidelity=infid, guard_penalty=guard, history=hist, lambda_history=lam,
                    adjoint_forcing=forc, iters_fwd=itf, iters_adj=ita, iters_term=itt)

    def discrete_adjoint_device(self, d_pcof_ptr, n_batch, d_target_ptr, order, d_grad_ptr, d_infid_ptr, d_guard_ptr,
                                stream_ptr=None):
        """Everything already in HBM (raw device pointers as ints); asynchronous on `stream_ptr`."""
        _check(lib().qgd_discrete_adjoint_device(self._h, C.c_void_p(d_pcof_ptr), int(n_batch), C.c_void_p(d_target_ptr),
                                                 int(order), C.c_void_p(d_grad_ptr), C.c_void_p(d_infid_ptr),
                                                 C.c_void_p(d_guard_ptr), C.c_void_p(stream_ptr or 0)))

    def adjoint_phase1(self, pcof, order):
        pc = self._pcof(pcof, self.P)
        B = pc.shape[1]
        final = np.zeros((self.N2, self.ncol, B), order="F")
        guard = np.zeros(B)
        _check(lib().qgd_adjoint_phase1(self._h, _dp(pc), B, int(order), _dp(final), _dp(guard)))
        return final, guard

    def adjoint_phase2(self, target_real, final_all):
        tgt = np.asfortranarray(target_real, dtype=np.float64)
        fa = np.asfortranarray(final_all, dtype=np.float64)
        B = fa.shape[2]
        grad = np.zeros((self.P, B), order="F")
        infid = np.zeros(B)
        _check(lib().qgd_adjoint_phase2(self._h, _dp(tgt), _dp(fa), _dp(grad), _dp(infid)))
        return grad, infid

    def eval_controls(self, pcof, times, nderiv, want_grad=False):
        times = np.ascontiguousarray(times, dtype=np.float64)
        pc = np.ascontiguousarray(pcof, dtype=np.float64)
        nt = times.size
        p = np.zeros((nderiv, self.Nc, nt), order="F")
        q = np.zeros((nderiv, self.Nc, nt), order="F")
        gp = np.zeros((self.P, nderiv, nt), order="F") if want_grad else None
        gq = np.zeros((self.P, nderiv, nt), order="F") if want_grad else None
        _check(lib().qgd_eval_controls(self._h, _dp(pc), _dp(times), nt, int(nderiv), _dp(p), _dp(q), _dp(gp), _dp(gq)))
        return p, q, gp, gq

    def compute_derivatives(self, uv, order, cre, cim, adjoint=False):
        uv = np.asfortranarray(uv, dtype=np.float64).copy(order="F")
        if uv.ndim == 2:
            uv = uv[:, :, None]
        cre = np.asfortranarray(cre, dtype=np.float64)
        cim = np.asfortranarray(cim, dtype=np.float64)
        _check(lib().qgd_compute_derivatives(self._h, _dp(uv), uv.shape[2], int(order), _dp(cre), _dp(cim), int(adjoint)))
        return uv

    def stats(self):
        s = _abi.qgd_stats_t()
        _check(lib().qgd_get_stats(self._h, C.byref(s)))
        return {f: getattr(s, f) for f, _ in _abi.qgd_stats_t._fields_}


def comm_unique_id() -> bytes:
    """ncclGetUniqueId through the C ABI (rank 0 calls it and distributes the 128 bytes)."""
    buf = C.create_string_buffer(128)
    _check(lib().qgd_comm_get_unique_id(buf))
    return buf.raw


class MultiGPU:
    """qgd_init_multi_gpu: ONE process driving n GPUs (per-device handles + an NCCL communicator inside the library)."""

    def __init__(self, prob, controls, n_gpus, devices=None):
        self._pack = _abi.ProblemPack(prob, controls)
        self._mg = C.c_void_p()
        dev = (C.c_int32 * n_gpus)(*devices) if devices is not None else None
        _check(lib().qgd_init_multi_gpu(self._pack.ref(), int(n_gpus), dev, C.byref(self._mg)))
        self.n_gpus = int(n_gpus)
        self.P = self._pack.n_coeff
        self.N2 = prob.real_system_size
        self.nic = prob.N_initial_conditions

    def set_option(self, key, value):
        for i in range(self.n_gpus):
            h = C.c_void_p()
            _check(lib().qgd_multi_handle(self._mg, i, C.byref(h)))
            _check(lib().qgd_set_option(h, int(key), int(value)))

    def set_nsteps(self, nsteps):
        _check(lib().qgd_multi_set_nsteps(self._mg, int(nsteps)))

    def set_gmres_tolerances(self, abstol, reltol):
        _check(lib().qgd_multi_set_gmres_tolerances(self._mg, float(abstol), float(reltol)))

    def discrete_adjoint(self, pcof, target_real, order=2, shard=SHARD_COLUMNS):
        pc = Handle._pcof(pcof, self.P)
        B = pc.shape[1]
        tgt = np.asfortranarray(target_real, dtype=np.float64)
        if tgt.shape != (self.N2, self.nic):
            raise ValueError(f"target must be the real-stacked [2N, nic] = {(self.N2, self.nic)} array, got {tgt.shape}")
        grad = np.zeros((self.P, B), order="F")
        infid = np.zeros(B)
        guard = np.zeros(B)
        _check(lib().qgd_multi_discrete_adjoint(self._mg, _dp(pc), B, _dp(tgt), int(order), int(shard), _dp(grad), _dp(infid),
                                                _dp(guard)))
        return dict(grad=grad, infidelity=infid, guard_penalty=guard)

    def close(self):
        if getattr(self, "_mg", None) is not None and self._mg.value:
            lib().qgd_multi_destroy(self._mg)
            self._mg = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def problem_key(prob, controls):
    """Fingerprint of the immutable parts of (prob, controls): shapes, physics parameters and a content hash of the
    operators / initial conditions / control descriptors.  nsteps and the GMRES tolerances are mutable knobs
    (examples/cnot3_optimize_gate.jl:51-55) and are re-synced on every call instead."""
    import hashlib

    import scipy.sparse as sp

    from .controls import as_control_list

    hsh = hashlib.blake2b(digest_size=16)

    def feed(a):
        if sp.issparse(a):
            a = a.tocsc()
            for part in (a.indptr, a.indices, a.data):
                hsh.update(np.ascontiguousarray(part).tobytes())
        else:
            hsh.update(np.ascontiguousarray(np.asarray(a, dtype=np.float64)).tobytes())

    for a in (prob.system_sym, prob.system_asym, *prob.sym_operators, *prob.asym_operators, prob.u0, prob.v0,
              prob.guard_subspace_projector):
        feed(a)
    ctl = tuple(_control_descriptor(c) for c in as_control_list(controls))
    return (prob.N_tot_levels, prob.N_ess_levels, prob.N_initial_conditions, prob.N_operators, float(prob.tf),
            int(prob.preconditioner_type), ctl, hsh.hexdigest())


def _control_descriptor(c):
    """Everything qgd_create reads from a control (type and shape parameters; coefficient values are not part of it)."""
    fields = []
    for name in ("N_coeff", "tf", "N_amplitudes", "D1", "degree", "N_basis_functions"):
        if hasattr(c, name):
            fields.append((name, float(getattr(c, name))))
    base = getattr(c, "base_control", None)
    if base is not None:
        fields.append(("carrier", _control_descriptor(base), tuple(float(x) for x in c.carrier_frequencies)))
    return (type(c).__name__, tuple(fields))


# Device handles of the api functions, keyed on the CONTENT fingerprint above (never on id(): CPython reuses the ids
# of freed objects, and a temporary problem must not inherit the operators of a dead one).  Least recently used
# handles are closed beyond `HANDLE_CACHE_SIZE`, which also bounds the device memory the cache can pin.
HANDLE_CACHE_SIZE = 4
_HANDLES = {}  # key -> Handle, insertion order = recency


def get_handle(prob, controls, device: int = -1) -> Handle:
    """Cached device handle for (prob, controls); follows the reference's in-place mutation of
    prob.nsteps / prob.gmres_abstol / prob.gmres_reltol (examples/cnot3_optimize_gate.jl:51-55)."""
    key = (problem_key(prob, controls), device)
    h = _HANDLES.pop(key, None)
    if h is None:
        h = Handle(prob, controls, device)
        h._tol = (prob.gmres_abstol, prob.gmres_reltol)
    _HANDLES[key] = h  # most recently used last
    while len(_HANDLES) > HANDLE_CACHE_SIZE:
        _, old = next(iter(_HANDLES.items()))
        _HANDLES.pop(next(iter(_HANDLES)))
        old.close()
    if h.nsteps != prob.nsteps:
        h.set_nsteps(prob.nsteps)
    if h._tol != (prob.gmres_abstol, prob.gmres_reltol):
        h.set_gmres_tolerances(prob.gmres_abstol, prob.gmres_reltol)
        h._tol = (prob.gmres_abstol, prob.gmres_reltol)
    return h


def clear_handles():
    for h in _HANDLES.values():
        h.close()
    _HANDLES.clear()
    from . import convergence  # the per-level handles of get_histories

    convergence.clear_level_pool()
