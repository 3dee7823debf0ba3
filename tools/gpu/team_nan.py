import sys
import numpy as np
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import __graft_entry__ as g
q = g.load_package()
rabi = q.construct_rabi_prob(tf=np.pi, gmres_abstol=1e-15, gmres_reltol=1e-15, nsteps=10)
ctl = q.CarrierControl(q.FortranBSplineControl(16, 20, rabi.tf), [-10, -1, 0, 1, 10])
rng = np.random.default_rng(0)
pcof = rng.random(ctl.N_coeff); target = rng.random((2, 2)) + 1j * rng.random((2, 2))
rng = np.random.default_rng(11)
pcofs = np.stack([pcof, pcof * (1.0 + 0.3 * rng.standard_normal(len(pcof)))], axis=1)
h = q.Handle(rabi, ctl)
for team in (1, 2):
    h.set_option(q.backend.OPT_LATENCY_TEAM, team)
    out = h.discrete_adjoint(pcofs, q.complex_to_real(target), order=8, want_iters=True)
    print("team", team, "nan:", np.isnan(out["grad"]).any(axis=0), "iters fwd max", out["iters_fwd"].max(axis=(0, 1)), "adj", out["iters_adj"].max(axis=(0, 1)),
          "infid", out["infidelity"])
