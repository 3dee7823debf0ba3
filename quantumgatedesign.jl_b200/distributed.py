"""One process per GPU: sharding of the gradient evaluation across the ranks of a torch.distributed group.

Two natural partitions:

* **control vectors** -- evaluations for different ``pcof`` are independent (the batched random-pcof sweep,
  examples/optimization_with_random_pcof.jl style).  The batch is split into contiguous blocks, there is no
  data-path collective; results are gathered at the end (`PcofShardedEvaluator`).
* **initial-condition columns** -- the columns of one evaluation are independent through both sweeps
  (`Threads.@threads for initial_condition_index`, src/forward_evolution.jl:48,332) and couple only through the
  global scalars <psi_N, R>, <psi_N, T> of the terminal condition (src/eval_grad_discrete_adjoint.jl:27-28) and
  the final sums.  Each rank owns a contiguous block of columns: forward sweep, all-gather of the final states
  (2N x nic doubles per control vector), backward sweep, all-reduce of [grad; guard] (`ColumnShardedEvaluator`).

`attach_library_communicator` is the product path for columns: it hands every rank's `backend.Handle` an NCCL
communicator INSIDE libqgd_b200.so (qgd_comm_init_rank); `Handle.discrete_adjoint` / `discrete_adjoint_device` then run
both exchanges on the device, on the sweep stream, and torch.distributed only distributes the 128-byte NCCL id.
`ColumnShardedEvaluator` (round 1: host-staged collectives around the two-phase ABI) stays as the reference
implementation of the exchange pattern the CPU tests exercise.

The evaluators only need a *backend* with the two-phase interface of the C ABI
(`adjoint_phase1(pcofs, order) -> (final_local, guard_local)`,
`adjoint_phase2(target_real, final_all) -> (grad_local, infidelity)`, `set_column_shard(begin, count)`,
`discrete_adjoint(pcofs, target_real, order=...) -> dict`), which `backend.Handle` provides on the GPU; the
CPU tests (gloo, world_size 2) plug in a stand-in with the same interface.
"""
from __future__ import annotations

import numpy as np


def partition(n: int, world: int, rank: int):
    """Contiguous block partition of range(n): -> (begin, count).  The first n % world ranks get one extra."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad rank / world size")
    base, rem = divmod(n, world)
    count = base + (1 if rank < rem else 0)
    begin = rank * base + min(rank, rem)
    return begin, count


def _dist():
    import torch.distributed as dist

    return dist


def _to_tensor(a, device):
    import torch

    t = torch.from_numpy(np.ascontiguousarray(a))
    return t.to(device) if device is not None else t


class ColumnShardedEvaluator:
    """discrete_adjoint with the initial-condition columns sharded over the ranks of `group`.

    Every rank passes the SAME pcofs [P, B] and target and receives the same (grad [P, B], infidelity [B],
    guard_penalty [B]).  `device` is the torch device the collectives run on (a CUDA device for NCCL, None for
    gloo)."""

    def __init__(self, backend, nic: int, group=None, device=None):
        dist = _dist()
        self.backend, self.group, self.device = backend, group, device
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self.nic = nic
        if self.world > nic:
            raise ValueError(f"cannot shard {nic} columns over {self.world} ranks")
        self.begin, self.count = partition(nic, self.world, self.rank)
        backend.set_column_shard(self.begin, self.count)

    def discrete_adjoint(self, pcofs, target_real, order=2):
        dist = _dist()
        pcofs = np.asarray(pcofs, dtype=np.float64)
        if pcofs.ndim == 1:
            pcofs = pcofs[:, None]
        P, B = pcofs.shape
        final_local, guard_local = self.backend.adjoint_phase1(pcofs, order)   # [2N, count, B], [B]
        N2 = final_local.shape[0]
        # all-gather of the final states; column blocks may differ by one, so pad to the largest block
        maxc = partition(self.nic, self.world, 0)[1]
        send = np.zeros((B, maxc, N2))
        send[:, : self.count, :] = np.transpose(final_local, (2, 1, 0))
        t_send = _to_tensor(send, self.device)
        t_all = [t_send.new_empty(t_send.shape) for _ in range(self.world)]
        dist.all_gather(t_all, t_send, group=self.group)
        final_all = np.zeros((N2, self.nic, B), order="F")
        for r, t in enumerate(t_all):
            b, c = partition(self.nic, self.world, r)
            final_all[:, b:b + c, :] = np.transpose(t.cpu().numpy()[:, :c, :], (2, 1, 0))
        grad_local, infid = self.backend.adjoint_phase2(target_real, final_all)  # [P, B], [B]
        red = _to_tensor(np.concatenate([np.asarray(grad_local).ravel(order="F"), np.asarray(guard_local)]), self.device)
        dist.all_reduce(red, group=self.group)
        red = red.cpu().numpy()
        grad = red[: P * B].reshape((P, B), order="F")
        guard = red[P * B:]
        return dict(grad=grad, infidelity=np.asarray(infid), guard_penalty=guard)


class PcofShardedEvaluator:
    """A batch of control vectors split over the ranks; no collective on the data path.  `discrete_adjoint`
    returns this rank's block; `gather` assembles the whole batch on every rank (for the caller that needs it)."""

    def __init__(self, backend, group=None, device=None):
        dist = _dist()
        self.backend, self.group, self.device = backend, group, device
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)

    def local_block(self, B: int):
        return partition(B, self.world, self.rank)

    def discrete_adjoint(self, pcofs_all, target_real, order=2):
        pcofs_all = np.asarray(pcofs_all, dtype=np.float64)
        b, c = self.local_block(pcofs_all.shape[1])
        if c == 0:
            P = pcofs_all.shape[0]
            return dict(grad=np.zeros((P, 0)), infidelity=np.zeros(0), guard_penalty=np.zeros(0))
        out = self.backend.discrete_adjoint(np.asfortranarray(pcofs_all[:, b:b + c]), target_real, order=order)
        return dict(grad=out["grad"], infidelity=out["infidelity"], guard_penalty=out["guard_penalty"])

    def gather(self, local, B: int):
        dist = _dist()
        P = local["grad"].shape[0]
        maxc = partition(B, self.world, 0)[1]
        b, c = self.local_block(B)
        send = np.zeros((maxc, P + 2))
        send[:c, :P] = local["grad"].T
        send[:c, P] = local["infidelity"]
        send[:c, P + 1] = local["guard_penalty"]
        t_send = _to_tensor(send, self.device)
        t_all = [t_send.new_empty(t_send.shape) for _ in range(self.world)]
        dist.all_gather(t_all, t_send, group=self.group)
        grad = np.zeros((P, B), order="F"); infid = np.zeros(B); guard = np.zeros(B)
        for r, t in enumerate(t_all):
            rb, rc = partition(B, self.world, r)
            a = t.cpu().numpy()[:rc]
            grad[:, rb:rb + rc] = a[:, :P].T
            infid[rb:rb + rc] = a[:, P]
            guard[rb:rb + rc] = a[:, P + 1]
        return dict(grad=grad, infidelity=infid, guard_penalty=guard)


def attach_library_communicator(handle, group=None, get_unique_id=None):
    """Column sharding with the collectives inside the library: rank 0 of `group` draws an NCCL unique id through the C
    ABI (qgd_comm_get_unique_id), torch.distributed broadcasts the 128 bytes, every rank attaches the communicator to
    its handle (qgd_comm_init_rank, collective), which also assigns the rank its column block.  Afterwards
    `handle.discrete_adjoint(pcofs, target, order=...)` on every rank (same arguments) returns the complete gradient."""
    dist = _dist()
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    if get_unique_id is None:
        from .backend import comm_unique_id as get_unique_id
    box = [get_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(box, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
    uid = box[0]
    if not isinstance(uid, (bytes, bytearray)) or len(uid) != 128:
        raise RuntimeError("NCCL unique id did not arrive")
    handle.comm_init_rank(world, rank, bytes(uid))
    return world, rank
