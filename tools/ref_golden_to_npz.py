#!/usr/bin/env python
"""Convert the raw outputs of tools/make_reference_golden.jl into tests/golden/ref_<case>.npz (see
tools/export_reference_inputs.py for the whole recipe)."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    src = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "tests", "golden", "ref_inputs")
    for name in sorted(os.listdir(src)):
        d = os.path.join(src, name)
        if not os.path.exists(os.path.join(d, "out_meta.json")):
            continue
        meta = json.load(open(os.path.join(d, "meta.json")))
        om = json.load(open(os.path.join(d, "out_meta.json")))
        N2, nic, nsteps = 2 * meta["N_tot_levels"], meta["N_initial_conditions"], meta["nsteps"]
        rd = lambda fn: np.fromfile(os.path.join(d, fn + ".f64"))  # noqa: E731
        np.savez_compressed(
            os.path.join(ROOT, "tests", "golden", f"ref_{name}.npz"), order=meta["order"], grad=rd("out_grad"),
            final_state=rd("out_final_state").reshape((N2, nic), order="F"),
            lambda0=rd("out_lambda0").reshape((N2, nsteps + 1, nic), order="F"), infidelity=om["infidelity"],
            guard_penalty=om["guard_penalty"], iters_fwd_total=om["ldiv_fwd"] - om["solves_fwd"],
            iters_adj_total=om["ldiv_adj"] - om["solves_adj"], julia=om["julia"], threads=om["threads"])
        print("wrote tests/golden/ref_%s.npz" % name)


if __name__ == "__main__":
    main()
