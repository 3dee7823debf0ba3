"""ctypes mirror of include/qgd_b200.h (struct layouts + marshalling of host objects).

Pure data plumbing: turns a `SchrodingerProb` + control collection into a `qgd_problem_t`
whose pointers reference numpy arrays kept alive by the returned `ProblemPack`.
The layouts follow the reference's own storage so a Julia `SparseMatrixCSC` /
`Matrix{Float64}` could be passed the same way: column-major Float64,
1-based Int64 `colptr` / `rowval`.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

try:  # scipy is only needed when sparse operators are used
    import scipy.sparse as sp
except Exception:  # pragma: no cover
    sp = None

c_double_p = C.POINTER(C.c_double)
c_int64_p = C.POINTER(C.c_int64)

QGD_MAT_DENSE = 0
QGD_MAT_CSC = 1
QGD_CONTROL_GRAPE = 1
QGD_CONTROL_BSPLINE2 = 2
QGD_CONTROL_FORTRAN_BSPLINE = 3
QGD_CONTROL_HOST_TABLE = 4
QGD_PRECOND_IDENTITY = 0
QGD_PRECOND_LU = 1
QGD_PRECOND_DIAGONAL = 2


class qgd_matrix_t(C.Structure):
    _fields_ = [
        ("kind", C.c_int32),
        ("reserved", C.c_int32),
        ("nrows", C.c_int64),
        ("ncols", C.c_int64),
        ("dense", c_double_p),
        ("nnz", C.c_int64),
        ("colptr", c_int64_p),
        ("rowval", c_int64_p),
        ("nzval", c_double_p),
    ]


class qgd_control_t(C.Structure):
    _fields_ = [
        ("type", C.c_int32),
        ("reserved", C.c_int32),
        ("tf", C.c_double),
        ("n_amplitudes", C.c_int64),
        ("D1", C.c_int64),
        ("degree", C.c_int64),
        ("n_basis", C.c_int64),
        ("n_carriers", C.c_int64),
        ("carrier_freqs", c_double_p),
    ]


class qgd_problem_t(C.Structure):
    _fields_ = [
        ("N_tot_levels", C.c_int64),
        ("N_ess_levels", C.c_int64),
        ("N_initial_conditions", C.c_int64),
        ("N_operators", C.c_int64),
        ("system_sym", qgd_matrix_t),
        ("system_asym", qgd_matrix_t),
        ("sym_operators", C.POINTER(qgd_matrix_t)),
        ("asym_operators", C.POINTER(qgd_matrix_t)),
        ("u0", c_double_p),
        ("v0", c_double_p),
        ("guard_subspace_projector", qgd_matrix_t),
        ("tf", C.c_double),
        ("nsteps", C.c_int64),
        ("gmres_abstol", C.c_double),
        ("gmres_reltol", C.c_double),
        ("preconditioner", C.c_int32),
        ("reserved", C.c_int32),
        ("controls", C.POINTER(qgd_control_t)),
    ]


class qgd_stats_t(C.Structure):
    _fields_ = [
        ("kernel_launches", C.c_int64),
        ("h2d_bytes", C.c_int64),
        ("d2h_bytes", C.c_int64),
        ("last_forward_ms", C.c_double),
        ("last_backward_ms", C.c_double),
        ("last_total_ms", C.c_double),
        ("fast_path_launches", C.c_int64),
        ("collectives", C.c_int64),
    ]


def dptr(a: np.ndarray):
    return a.ctypes.data_as(c_double_p)


def iptr(a: np.ndarray):
    return a.ctypes.data_as(c_int64_p)


def fvec(x, shape=None) -> np.ndarray:
    """Column-major float64 copy/view suitable for passing across the ABI."""
    a = np.asarray(x, dtype=np.float64)
    if shape is not None:
        a = a.reshape(shape, order="F") if a.ndim == 1 and len(shape) > 1 else a
    return np.asfortranarray(a)


class ProblemPack:
    """Owns every buffer referenced by `self.struct` (a qgd_problem_t)."""

    def __init__(self, prob, controls):
        from .controls import as_control_list, control_descriptor

        self._keep = []
        controls = as_control_list(controls)
        if len(controls) != prob.N_operators:
            raise ValueError(
                f"Number of controls {len(controls)} does not match number of control operators {prob.N_operators}."
            )
        s = qgd_problem_t()
        s.N_tot_levels = prob.N_tot_levels
        s.N_ess_levels = prob.N_ess_levels
        s.N_initial_conditions = prob.N_initial_conditions
        s.N_operators = prob.N_operators
        s.system_sym = self._mat(prob.system_sym)
        s.system_asym = self._mat(prob.system_asym)
        MatArr = qgd_matrix_t * max(prob.N_operators, 1)
        sym = MatArr(*[self._mat(m) for m in prob.sym_operators])
        asym = MatArr(*[self._mat(m) for m in prob.asym_operators])
        self._keep += [sym, asym]
        s.sym_operators = sym
        s.asym_operators = asym
        u0 = np.asfortranarray(np.asarray(prob.u0, dtype=np.float64).reshape(prob.N_tot_levels, -1))
        v0 = np.asfortranarray(np.asarray(prob.v0, dtype=np.float64).reshape(prob.N_tot_levels, -1))
        self._keep += [u0, v0]
        s.u0 = dptr(u0)
        s.v0 = dptr(v0)
        s.guard_subspace_projector = self._mat(prob.guard_subspace_projector)
        s.tf = float(prob.tf)
        s.nsteps = int(prob.nsteps)
        s.gmres_abstol = float(prob.gmres_abstol)
        s.gmres_reltol = float(prob.gmres_reltol)
        s.preconditioner = int(prob.preconditioner_type)
        CtlArr = qgd_control_t * max(len(controls), 1)
        descs = []
        for c in controls:
            d, keep = control_descriptor(c)
            self._keep.append(keep)
            descs.append(d)
        ctl = CtlArr(*descs)
        self._keep.append(ctl)
        s.controls = ctl
        self.struct = s
        self.n_coeff = sum(c.N_coeff for c in controls)

    def _mat(self, M) -> qgd_matrix_t:
        m = qgd_matrix_t()
        if sp is not None and sp.issparse(M):
            A = sp.csc_matrix(M, dtype=np.float64)
            A.sort_indices()
            colptr = (A.indptr.astype(np.int64) + 1).copy()
            rowval = (A.indices.astype(np.int64) + 1).copy()
            nz = np.ascontiguousarray(A.data, dtype=np.float64)
            self._keep += [colptr, rowval, nz]
            m.kind = QGD_MAT_CSC
            m.nrows, m.ncols = A.shape
            m.nnz = int(A.nnz)
            m.colptr = iptr(colptr)
            m.rowval = iptr(rowval)
            m.nzval = dptr(nz)
        else:
            A = np.asfortranarray(np.asarray(M, dtype=np.float64))
            self._keep.append(A)
            m.kind = QGD_MAT_DENSE
            m.nrows, m.ncols = A.shape
            m.dense = dptr(A)
            m.nnz = 0
        return m

    def ref(self):
        return C.byref(self.struct)
