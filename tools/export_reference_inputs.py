#!/usr/bin/env python
"""Export the inputs of the named parity cases for the REAL reference (QuantumGateDesign.jl, Julia), so that a machine
with Julia can pin this repository's oracle and CUDA path to the reference's own outputs:

    python tools/export_reference_inputs.py [outdir=tests/golden/ref_inputs]       # here (or anywhere with numpy)
    julia --threads 8 tools/make_reference_golden.jl tests/golden/ref_inputs          # on a machine with Julia + the reference
    python tools/ref_golden_to_npz.py tests/golden/ref_inputs                         # -> tests/golden/ref_<case>.npz
    python -m pytest tests/test_golden.py                                            # the ref_* fixtures are picked up

Neither Julia nor the reference can run in the build image or on the GPU box, so the ref_*.npz
fixtures are NOT part of this repository yet; tests/test_golden.py activates the comparisons when they exist.

Per case a directory <outdir>/<case>/ with meta.json and raw little-endian Float64 column-major arrays (*.f64): what
Julia reads with `read!(io, Array{Float64}(undef, dims...))`.  Random inputs come from numpy's default_rng (Julia's
MersenneTwister streams are not reproducible across languages), which is exactly why the inputs are exported."""
import json
import os
import sys

import numpy as np
import scipy.sparse as sp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from __graft_entry__ import load_package  # noqa: E402


def dense(a):
    return np.asfortranarray(a.toarray() if sp.issparse(a) else np.asarray(a), dtype=np.float64)


def control_meta(q, c):
    if isinstance(c, q.CarrierControl):
        d = control_meta(q, c.base_control)
        d["carrier_frequencies"] = [float(w) for w in c.carrier_frequencies]
        return d
    if isinstance(c, q.GRAPEControl):
        return {"type": "GRAPEControl", "N_amplitudes": c.N_amplitudes, "tf": c.tf}
    if isinstance(c, q.BSpline2Control):
        return {"type": "BSpline2Control", "D1": c.D1, "tf": c.tf}
    if isinstance(c, q.FortranBSplineControl):
        return {"type": "FortranBSplineControl", "degree": c.degree, "N_basis_functions": c.N_basis_functions, "tf": c.tf}
    raise ValueError(f"no reference constructor mapping for {type(c).__name__}")


def cases(q):
    import importlib.util

    spec = importlib.util.spec_from_file_location("make_golden", os.path.join(ROOT, "tests", "golden", "make_golden.py"))
    mg = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mg)
    c = dict(mg.cases(q))
    c["c2_cnot3_full"] = q.configs.cnot3(nsteps=550, tf=550.0, gmres_tol=1e-14)   # tests/test_gpu_parity.py::test_full_cnot3_order8
    return c


def main():
    q = load_package()
    out = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "tests", "golden", "ref_inputs")
    for name, (prob, controls, pcof, target, order) in cases(q).items():
        d = os.path.join(out, name)
        os.makedirs(d, exist_ok=True)

        def put(fn, a):
            np.asfortranarray(a, dtype=np.float64).ravel(order="F").tofile(os.path.join(d, fn + ".f64"))

        put("system_sym", dense(prob.system_sym)); put("system_asym", dense(prob.system_asym))
        for k in range(prob.N_operators):
            put(f"sym_op_{k + 1}", dense(prob.sym_operators[k])); put(f"asym_op_{k + 1}", dense(prob.asym_operators[k]))
        put("u0", prob.u0); put("v0", prob.v0); put("guard_subspace_projector", dense(prob.guard_subspace_projector))
        put("pcof", pcof); put("target_re", np.real(target)); put("target_im", np.imag(target))
        meta = {
            "name": name, "order": int(order), "N_tot_levels": prob.N_tot_levels, "N_ess_levels": prob.N_ess_levels,
            "N_initial_conditions": prob.N_initial_conditions, "N_operators": prob.N_operators, "tf": float(prob.tf),
            "nsteps": int(prob.nsteps), "gmres_abstol": prob.gmres_abstol, "gmres_reltol": prob.gmres_reltol,
            "preconditioner": ["IdentityPreconditioner", "LUPreconditioner", "DiagonalHamiltonianPreconditioner"][int(prob.preconditioner_type)], "sparse": bool(sp.issparse(prob.system_sym)),
            "N_coeff": int(len(pcof)), "controls": [control_meta(q, c) for c in q.as_control_list(controls)],
        }
        json.dump(meta, open(os.path.join(d, "meta.json"), "w"), indent=1)
        print("exported", name, "->", d)


if __name__ == "__main__":
    main()
