"""CPU experiment (oracle only, test infrastructure): how far do GMRES iteration counts and the gradient move from the
strict modified Gram-Schmidt of the reference when the orthogonalisation is classical inside blocks of B basis vectors
(modified across blocks)?  Run:  python tools/gs_block_experiment.py [nsteps] [tol ...]
Each block width runs in its own process (the oracle reads QGDO_GS_BLOCK once)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CHILD = r"""
import sys, os, json
sys.path.insert(0, %r); sys.path.insert(0, os.path.join(%r, "oracle"))
import numpy as np
import __graft_entry__ as G
q = G.load_package()
import oracle as O
nsteps, tol = int(sys.argv[1]), float(sys.argv[2])
prob, controls, pcof, target, order = q.configs.cnot3(nsteps=nsteps, tf=float(nsteps), gmres_tol=tol)
r = O.discrete_adjoint(prob, controls, pcof, target, order=order)
np.savez(sys.argv[3], grad=r["grad"], itf=r["iters_fwd"], ita=r["iters_adj"], infid=r["infidelity"])
""" % (ROOT, ROOT)


def run(block, nsteps, tol, out):
    env = dict(os.environ, QGDO_GS_BLOCK=str(block))
    subprocess.check_call([sys.executable, "-c", CHILD, str(nsteps), repr(tol), out], env=env)


if __name__ == "__main__":
    import numpy as np
    nsteps = int(sys.argv[1]) if len(sys.argv) > 1 else 60
    tols = [float(t) for t in sys.argv[2:]] or [1e-12, 1e-14]
    for tol in tols:
        run(0, nsteps, tol, "/tmp/gs_ref.npz")
        ref = np.load("/tmp/gs_ref.npz")
        nsolves = ref["itf"].size + ref["ita"].size
        print(f"tol {tol:g}: strict MGS  total iterations {int(ref['itf'].sum() + ref['ita'].sum())} in {nsolves} solves")
        for block in (8, 16, 32, 64, 128):
            run(block, nsteps, tol, "/tmp/gs_blk.npz")
            r = np.load("/tmp/gs_blk.npz")
            d = np.concatenate([(r["itf"] - ref["itf"]).ravel(), (r["ita"] - ref["ita"]).ravel()])
            gerr = np.abs(r["grad"] - ref["grad"]).max() / np.abs(ref["grad"]).max()
            print(f"  block {block:3d}: solves with a different count {int((d != 0).sum())} (max |diff| {int(np.abs(d).max())}), "
                  f"total iteration diff {int(d.sum())}, gradient rel diff {gerr:.2e}, infidelity diff {abs(float(r['infid']) - float(ref['infid'])):.2e}")
