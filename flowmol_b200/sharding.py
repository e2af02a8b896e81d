"""Molecule-level sharding of one batch across the GPUs of a box.

Molecules never interact (every reduction of the path is per molecule), so a batch is cut into `world` CONTIGUOUS ranges of
global molecule ids balanced by cost ~ n(n-1) (directed edges); each rank runs the whole trajectory on its range with no
collective inside a timestep, and the results are gathered once at the end.  Noise is keyed by the GLOBAL molecule id
(`mol_id_offset` = first id of the range), so any world size returns bit-identical molecules.
"""
import numpy as np
import torch


def partition(n_atoms, world):
    """-> list of (lo, hi) global molecule ranges, one per rank, contiguous, balanced by sum n(n-1) (greedy prefix cut)."""
    n = np.asarray(n_atoms, dtype=np.int64)
    cost = np.concatenate([[0], np.cumsum(n * (n - 1) + 64)])       # +64: per-molecule tile padding / fixed overhead
    total = cost[-1]
    cuts = [0]
    for r in range(1, world):
        target = total * r / world
        k = int(np.searchsorted(cost, target, side="left"))
        k = min(max(k, cuts[-1] + (1 if len(n) - cuts[-1] > world - r else 0)), len(n) - (world - r))
        cuts.append(max(k, cuts[-1]))
    cuts.append(len(n))
    return [(cuts[r], cuts[r + 1]) for r in range(world)]


def pack_result(out, n_max, u_max):
    """x f32[N,3] | a u8[N] | c u8[N] | e u8[U]  ->  one uint8 buffer padded to (n_max, u_max) for a fixed-size gather."""
    N, U = out["x"].shape[0], out["e"].shape[0]
    buf = torch.zeros(n_max * 14 + u_max, dtype=torch.uint8, device=out["x"].device)
    buf[:N * 12] = out["x"].contiguous().view(torch.uint8).reshape(-1)
    buf[n_max * 12:n_max * 12 + N] = out["a"].to(torch.uint8)
    buf[n_max * 13:n_max * 13 + N] = out["c"].to(torch.uint8)
    buf[n_max * 14:n_max * 14 + U] = out["e"].to(torch.uint8)
    return buf


def unpack_result(buf, N, U, n_max):
    x = buf[:N * 12].clone().view(torch.float32).reshape(N, 3)
    return {"x": x, "a": buf[n_max * 12:n_max * 12 + N].clone(), "c": buf[n_max * 13:n_max * 13 + N].clone(),
            "e": buf[n_max * 14:n_max * 14 + U].clone()}


def gather_results(out, n_atoms, ranges, rank, world, group=None):
    """The path's single exchange step: every rank's (x_1, a_1, c_1, e_1) to rank 0 (NCCL over NVLink on GPUs, gloo in
    the CPU tests).  Returns the concatenated global result on rank 0, None elsewhere."""
    import torch.distributed as dist
    n = np.asarray(n_atoms, dtype=np.int64)
    Ns = [int(n[lo:hi].sum()) for lo, hi in ranges]
    Us = [int((n[lo:hi] * (n[lo:hi] - 1) // 2).sum()) for lo, hi in ranges]
    n_max, u_max = max(Ns), max(Us)
    send = pack_result(out, n_max, u_max)
    recv = [torch.zeros_like(send) for _ in range(world)] if rank == 0 else None
    dist.gather(send, recv, dst=0, group=group)
    if rank != 0:
        return None
    parts = [unpack_result(recv[r], Ns[r], Us[r], n_max) for r in range(world)]
    return {k: torch.cat([p[k] for p in parts]) for k in "xace"}
