"""Host-side mirror of the reference's control types on the hot path.

These are data carriers only: the values / time-derivatives / parameter-gradients are
evaluated on the GPU by the control-table kernel (csrc/qgd_controls.cuh, kernel K1).
Names, constructor arguments and the coefficient layout follow the reference:

* `GRAPEControl(N_amplitudes, tf)`           -- src/Controls/grape_control.jl:18-26
* `BSpline2Control(D1, tf)`                  -- src/Controls/bspline_control.jl:21-43
* `FortranBSplineControl(degree, N_basis_functions, tf)` -- src/Controls/FortranBSpline.jl:16-61
* `CarrierControl(base_control, carrier_frequencies)`    -- src/Controls/CarrierControl.jl:5-23

`pcof` is the concatenation of the per-control slices (src/Controls/Control.jl:67-96); inside
a slice the first half drives p, the second half q; a CarrierControl slice is one base slice
per carrier frequency (CarrierControl.jl:44-46).
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List, Sequence

import numpy as np

from . import _abi


class AbstractControl:
    """Protocol: fields `N_coeff`, `tf` (src/Controls/Control.jl:6-27)."""

    N_coeff: int
    tf: float

    # A single control behaves as a 1-element collection (Control.jl:44-54,167-172)
    def __len__(self):
        return 1

    def __getitem__(self, i):
        if i != 0:
            raise IndexError(i)
        return self

    def __iter__(self):
        yield self


@dataclass
class GRAPEControl(AbstractControl):
    N_amplitudes: int
    tf: float
    N_coeff: int = field(init=False)

    def __post_init__(self):
        self.N_amplitudes = int(self.N_amplitudes)
        self.tf = float(self.tf)
        if self.N_amplitudes < 1:
            raise ValueError("N_amplitudes must be >= 1")
        self.N_coeff = 2 * self.N_amplitudes


@dataclass
class BSpline2Control(AbstractControl):
    D1: int
    tf: float
    N_coeff: int = field(init=False)

    def __post_init__(self):
        self.D1 = int(self.D1)
        self.tf = float(self.tf)
        if self.D1 < 3:
            raise ValueError(f"Number of coefficients per spline (D1 = {self.D1}) must be >= 3.")
        self.N_coeff = 2 * self.D1


@dataclass
class FortranBSplineControl(AbstractControl):
    degree: int
    N_basis_functions: int
    tf: float
    N_coeff: int = field(init=False)

    def __post_init__(self):
        self.degree = int(self.degree)
        self.N_basis_functions = int(self.N_basis_functions)
        self.tf = float(self.tf)
        order = self.degree + 1
        n_knots = self.N_basis_functions + order
        n_distinct = n_knots - 2 * (order - 1)
        if n_distinct < 2:
            raise ValueError("FortranBSplineControl: too few basis functions for this degree")
        if order > 20:
            raise ValueError("FortranBSplineControl: pppack supports order <= 20 (src/Fortran/bsplvb.f:62)")
        self.N_coeff = 2 * self.N_basis_functions
        self.bspline_order = order
        self.N_knots = n_knots
        self.N_distinct_knots = n_distinct


class CarrierControl(AbstractControl):
    def __init__(self, base_control: AbstractControl, carrier_frequencies: Sequence[float]):
        if isinstance(base_control, CarrierControl):
            raise ValueError("CarrierControl of a CarrierControl is not supported")
        self.base_control = base_control
        self.carrier_frequencies = np.ascontiguousarray(carrier_frequencies, dtype=np.float64)
        if self.carrier_frequencies.ndim != 1 or self.carrier_frequencies.size < 1:
            raise ValueError("carrier_frequencies must be a non-empty vector")
        self.N_coeffs_per_frequency = base_control.N_coeff
        self.N_coeff = base_control.N_coeff * self.carrier_frequencies.size
        self.tf = base_control.tf

    def __repr__(self):
        return f"CarrierControl({self.base_control!r}, {self.carrier_frequencies.tolist()})"


class HostEvaluatedControl(AbstractControl):
    """Any other AbstractControl (src/Controls/Control.jl:6-27 -- GeneralBSplineControl, the Hermite controls,
    bcarrier2, user-defined ones): the device kernels do not know the family, so the HOST evaluates the control
    protocol and the sweeps consume tables (include/qgd_b200.h, qgd_*_tables).

    `derivatives(t, pcof_local, nderiv)` returns the un-scaled time derivatives of orders 0..nderiv-1,
    `(p[nderiv], q[nderiv], grad_p[nderiv, N_coeff], grad_q[nderiv, N_coeff])` -- what eval_p_derivative /
    eval_q_derivative / eval_grad_p_derivative! / eval_grad_q_derivative! return in the reference."""

    def __init__(self, N_coeff, tf, derivatives, linear=True):
        self.N_coeff = int(N_coeff)
        self.tf = float(tf)
        self.derivatives = derivatives
        self.linear = bool(linear)  # p, q linear in pcof: the gradient table is shared by a batch of control vectors

    def __repr__(self):
        return f"HostEvaluatedControl(N_coeff={self.N_coeff}, tf={self.tf})"


class SinCosControl(HostEvaluatedControl):
    """SinCosControl(tf; frequency) (src/Controls/sincos_control.jl:1-27): p = sin(w t) pcof[1], q = cos(w t) pcof[2];
    a family the device control kernels do not evaluate, served through the host-table path."""

    def __init__(self, tf, frequency=1.0):
        self.frequency = float(frequency)
        w = self.frequency

        def derivs(t, pc, nderiv):
            j = np.arange(nderiv)
            s = w ** j * np.sin(w * t + j * np.pi / 2)
            c = w ** j * np.cos(w * t + j * np.pi / 2)
            gp = np.zeros((nderiv, 2)); gq = np.zeros((nderiv, 2))
            gp[:, 0] = s; gq[:, 1] = c
            return s * pc[0], c * pc[1], gp, gq

        super().__init__(2, tf, derivs, linear=True)


def has_host_controls(controls) -> bool:
    return any(isinstance(c, HostEvaluatedControl) for c in as_control_list(controls))


def build_control_tables(controls, pcofs, tf, nsteps, m):
    """Tables of the qgd_*_tables entry points for host-evaluated controls (t_n = n tf / nsteps):
    cvals [Nc, 1+m, 2, 1+nsteps, B] = p_k^(j)/j!, q_k^(j)/j!;  table [P, 1+m, 2, 1+nsteps] = d/dtheta of the same."""
    from math import factorial

    cl = as_control_list(controls)
    for c in cl:
        if not isinstance(c, HostEvaluatedControl):
            raise TypeError("host control tables: every control of the collection must be a HostEvaluatedControl "
                            f"(got {type(c).__name__}; the device-evaluated families have no host evaluator here)")
    pcofs = np.asarray(pcofs, dtype=np.float64)
    if pcofs.ndim == 1:
        pcofs = pcofs[:, None]
    P, B = pcofs.shape
    if B > 1 and not all(c.linear for c in cl):
        raise ValueError("controls that are nonlinear in pcof need one call per control vector (the gradient table is shared)")
    Nc, Nt = len(cl), nsteps + 1
    cvals = np.zeros((Nc, m + 1, 2, Nt, B), order="F")
    table = np.zeros((P, m + 1, 2, Nt), order="F")
    inv_fact = np.array([1.0 / factorial(j) for j in range(m + 1)])
    sl = control_slices(cl)
    for n in range(Nt):
        t = n * tf / nsteps
        for k, c in enumerate(cl):
            a, b = sl[k]
            for ib in range(B):
                pv, qv, gp, gq = c.derivatives(t, pcofs[a:b, ib], m + 1)
                cvals[k, :, 0, n, ib] = np.asarray(pv) * inv_fact
                cvals[k, :, 1, n, ib] = np.asarray(qv) * inv_fact
                if ib == 0:
                    table[a:b, :, 0, n] = (np.asarray(gp) * inv_fact[:, None]).T
                    table[a:b, :, 1, n] = (np.asarray(gq) * inv_fact[:, None]).T
    return cvals, table


def as_control_list(controls) -> List[AbstractControl]:
    if isinstance(controls, AbstractControl):
        return [controls]
    return list(controls)


def get_number_of_control_parameters(controls) -> int:
    """src/Controls/Control.jl:94-96"""
    return sum(c.N_coeff for c in as_control_list(controls))


def control_slices(controls):
    """Start/stop of every control's slice of pcof (get_control_vector_slice, Control.jl:67-75)."""
    out, start = [], 0
    for c in as_control_list(controls):
        out.append((start, start + c.N_coeff))
        start += c.N_coeff
    return out


def control_descriptor(c: AbstractControl):
    """-> (qgd_control_t, keepalive)"""
    d = _abi.qgd_control_t()
    keep = None
    base = c
    if isinstance(c, CarrierControl):
        base = c.base_control
        keep = c.carrier_frequencies
        d.n_carriers = keep.size
        d.carrier_freqs = _abi.dptr(keep)
    else:
        d.n_carriers = 0
    d.tf = float(base.tf)
    if isinstance(base, HostEvaluatedControl):
        if isinstance(c, CarrierControl):
            raise TypeError("CarrierControl of a host-evaluated control: evaluate the carrier on the host as well")
        d.type = _abi.QGD_CONTROL_HOST_TABLE
        d.n_amplitudes = base.N_coeff
    elif isinstance(base, GRAPEControl):
        d.type = _abi.QGD_CONTROL_GRAPE
        d.n_amplitudes = base.N_amplitudes
    elif isinstance(base, BSpline2Control):
        d.type = _abi.QGD_CONTROL_BSPLINE2
        d.D1 = base.D1
    elif isinstance(base, FortranBSplineControl):
        d.type = _abi.QGD_CONTROL_FORTRAN_BSPLINE
        d.degree = base.degree
        d.n_basis = base.N_basis_functions
    else:
        raise TypeError(
            f"control type {type(base).__name__} is not evaluated on the device (GRAPE, BSpline2, FortranBSpline, "
            "Carrier are): wrap it in HostEvaluatedControl to use the host-table path"
        )
    return d, keep
