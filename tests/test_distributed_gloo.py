"""Host-side multi-GPU logic on CPU: world_size-2 gloo runs of the column-sharded and pcof-sharded evaluators
(quantumgatedesign.jl_b200/distributed.py) against the same evaluation done by one rank.  The GPU `Handle` is
replaced by a stand-in with the same two-phase interface whose arithmetic is separable over columns exactly like
the real path (final states per column, two global scalars coupling the columns, gradient = sum over columns)."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


class StandInBackend:
    """Same call sequence and array shapes as backend.Handle; cheap deterministic arithmetic."""

    def __init__(self, N2, nic, P):
        self.N2, self.nic, self.P = N2, nic, P
        self.b, self.c = 0, nic

    def set_column_shard(self, begin, count):
        assert 0 <= begin and begin + count <= self.nic and count >= 1
        self.b, self.c = begin, count

    def _final(self, col, pc):
        r = np.arange(self.N2)
        return np.sin(0.1 * (col + 1) * (r + 1) + pc[: self.N2 % len(pc) + 1].sum()) + 0.01 * pc.sum()

    def adjoint_phase1(self, pcofs, order):
        B = pcofs.shape[1]
        final = np.zeros((self.N2, self.c, B), order="F")
        guard = np.zeros(B)
        for bi in range(B):
            for j in range(self.c):
                final[:, j, bi] = self._final(self.b + j, pcofs[:, bi])
                guard[bi] += 1e-3 * (self.b + j + 1) * np.cos(pcofs[:, bi]).sum()
        self._pcofs = pcofs
        return final, guard

    def adjoint_phase2(self, target_real, final_all):
        assert final_all.shape[:2] == (self.N2, self.nic)
        B = final_all.shape[2]
        grad = np.zeros((self.P, B), order="F")
        infid = np.zeros(B)
        for bi in range(B):
            dR = float((final_all[:, :, bi] * target_real).sum())   # couples ALL columns, like <psi_N, R>
            infid[bi] = 1.0 - dR * dR
            for j in range(self.c):
                col = self.b + j
                grad[:, bi] += dR * np.cos(0.3 * (col + 1) * np.arange(1, self.P + 1)) * self._pcofs[:, bi]
        return grad, infid

    def discrete_adjoint(self, pcofs, target_real, order=2):
        keep = (self.b, self.c)
        self.set_column_shard(0, self.nic)
        final, guard = self.adjoint_phase1(np.asarray(pcofs), order)
        grad, infid = self.adjoint_phase2(target_real, final)
        self.set_column_shard(*keep)
        return dict(grad=grad, infidelity=infid, guard_penalty=guard)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, nic, B, q):
    import torch.distributed as dist

    sys.path.insert(0, ROOT)
    from __graft_entry__ import load_package

    pkg = load_package()
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        N2, P = 6, 7
        rng = np.random.default_rng(5)
        pcofs = np.asfortranarray(rng.standard_normal((P, B)))
        target = rng.standard_normal((N2, nic))
        ref = StandInBackend(N2, nic, P).discrete_adjoint(pcofs, target)
        # columns sharded
        ev = pkg.distributed.ColumnShardedEvaluator(StandInBackend(N2, nic, P), nic)
        out = ev.discrete_adjoint(pcofs, target, order=4)
        ok_cols = (np.allclose(out["grad"], ref["grad"], rtol=1e-13, atol=1e-13)
                   and np.allclose(out["infidelity"], ref["infidelity"], rtol=1e-13)
                   and np.allclose(out["guard_penalty"], ref["guard_penalty"], rtol=1e-13))
        # control vectors sharded
        pe = pkg.distributed.PcofShardedEvaluator(StandInBackend(N2, nic, P))
        loc = pe.discrete_adjoint(pcofs, target, order=4)
        b, c = pe.local_block(B)
        ok_local = np.allclose(loc["grad"], ref["grad"][:, b:b + c], rtol=1e-13, atol=1e-13)
        allp = pe.gather(loc, B)
        ok_pcof = (np.allclose(allp["grad"], ref["grad"], rtol=1e-13, atol=1e-13)
                   and np.allclose(allp["infidelity"], ref["infidelity"], rtol=1e-13)
                   and np.allclose(allp["guard_penalty"], ref["guard_penalty"], rtol=1e-13))
        # in-library communicator: rank 0's NCCL id must reach every rank, each rank attaches with its own rank number
        class _Handle:
            def comm_init_rank(self, n_ranks, rk, uid):
                self.args = (n_ranks, rk, uid)

        hh = _Handle()
        pkg.distributed.attach_library_communicator(hh, get_unique_id=lambda: bytes([rank + 1]) * 128)
        ok_comm = hh.args == (world, rank, bytes([1]) * 128)
        q.put((rank, ev.begin, ev.count, bool(ok_cols), bool(ok_local), bool(ok_pcof), bool(ok_comm)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("nic,B", [(8, 3), (5, 4), (2, 1)])
def test_two_rank_sharding_equals_single_rank(nic, B):
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, nic, B, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    res.sort()
    assert res[0][1] == 0 and res[0][1] + res[0][2] == res[1][1] and res[1][1] + res[1][2] == nic  # blocks tile the columns
    for r in res:
        assert r[3], f"column-sharded result differs on rank {r[0]}"
        assert r[4] and r[5], f"pcof-sharded result differs on rank {r[0]}"
        assert r[6], f"NCCL unique id of rank 0 did not reach rank {r[0]} / wrong rank number passed to the library"


def test_partition_tiles_the_range(q):
    part = q.distributed.partition
    for n in (1, 2, 5, 8, 64, 1024):
        for world in (1, 2, 3, 4, 8):
            blocks = [part(n, world, r) for r in range(world)]
            assert blocks[0][0] == 0
            for (b0, c0), (b1, _) in zip(blocks, blocks[1:]):
                assert b0 + c0 == b1
            assert blocks[-1][0] + blocks[-1][1] == n
            assert max(c for _, c in blocks) - min(c for _, c in blocks) <= 1
    with pytest.raises(ValueError):
        part(4, 2, 2)
