"""Multi-GPU entry point of the batched SCP solvers: shard -> solve -> gather (boundary item (b)-7).

The reference solves one trajectory per `solve(pbm)` call (ptr.jl:448); a batch of seeds is embarrassingly parallel, so
the B seeds are cut into contiguous blocks of ceil(B / world) seeds, one block per rank (= one process per GPU), every
rank runs the device-resident loop on its block (no collective on the data path), and ONE all_gather over
NCCL / NVLink brings the converged (xd, ud, p, status, iterations, cost, deviation, feas) of all seeds to every rank.
`torch.distributed` is the plumbing (process group, NCCL communicator); on a CPU-only box the same code runs over gloo,
which is what tests/test_multirank_gloo.py exercises.
"""
from __future__ import annotations

import numpy as np


def shard_bounds(B: int, world: int, rank: int):
    """[lo, hi) of the contiguous block of seeds owned by `rank`: blocks of ceil(B / world), the last ones may be short
    or empty (256 seeds -> 32 per GPU on 8 GPUs)."""
    per = -(-int(B) // int(world))
    lo = min(B, rank * per)
    return lo, min(B, lo + per)


def _pack(sol, N, nx, nu, np_):
    """per-seed rows [xd | ud | p | status | iterations | cost | deviation | feas] as float64 (integers are exact)"""
    n = sol.xd.shape[0]
    return np.concatenate([sol.xd.reshape(n, N * nx), sol.ud.reshape(n, N * nu), sol.p.reshape(n, np_),
                           sol.raw_status.reshape(n, 1).astype(np.float64), sol.iterations.reshape(n, 1).astype(np.float64),
                           sol.cost.reshape(n, 1), sol.deviation.reshape(n, 1), sol.feas.reshape(n, 1).astype(np.float64)],
                          axis=1)


def gather_rows(rows: np.ndarray, B: int, world: int, rank: int, dist=None, device=None):
    """all_gather of the per-rank row blocks (shape (hi-lo, width)) into the full (B, width) array on every rank.
    Blocks are padded to ceil(B / world) rows so that one fixed-size collective suffices."""
    per = -(-int(B) // int(world))
    width = rows.shape[1]
    if world == 1 or dist is None:
        return np.ascontiguousarray(rows)
    import torch
    buf = np.zeros((per, width))
    buf[:rows.shape[0]] = rows
    t = torch.from_numpy(buf)
    if device is not None:
        t = t.to(device)
    out = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(out, t)
    full = np.zeros((B, width))
    for r in range(world):
        lo, hi = shard_bounds(B, world, r)
        if hi > lo:
            full[lo:hi] = out[r][:hi - lo].cpu().numpy()
    return full


def solve_sharded(algo, pbm, guesses, dist=None, device=None, **cone_opts):
    """algo: the ptr / scvx / gusto module of this package; guesses = (xd0 (B,N,nx), ud0 (B,N,nu), p0 (B,np)) for the WHOLE
    batch, identical on every rank.  Returns (full-batch SCPBatchSolution on every rank, this rank's local solution).
    dist: torch.distributed (initialised) or None for a single process; device: torch device of the gather buffers
    (cuda:<local_rank> for NCCL, None for gloo)."""
    xd0, ud0, p0 = (np.ascontiguousarray(a, dtype=np.float64) for a in guesses)
    B = xd0.shape[0]
    world = dist.get_world_size() if dist is not None else 1
    rank = dist.get_rank() if dist is not None else 0
    lo, hi = shard_bounds(B, world, rank)
    N, nx, nu, np_ = pbm.pars.N, pbm.traj.nx, pbm.traj.nu, pbm.traj.np
    local = None
    if hi > lo:
        local = algo.solve(pbm, (xd0[lo:hi], ud0[lo:hi], p0[lo:hi]), **cone_opts)
        rows = _pack(local, N, nx, nu, np_)
    else:
        rows = np.zeros((0, N * nx + N * nu + np_ + 5))
    full = gather_rows(rows, B, world, rank, dist, device)
    o = 0
    xd = full[:, o:o + N * nx].reshape(B, N, nx); o += N * nx
    ud = full[:, o:o + N * nu].reshape(B, N, nu); o += N * nu
    p = full[:, o:o + np_]; o += np_
    raw = full[:, o].astype(np.int32); its = full[:, o + 1].astype(np.int32)
    cost, dev, feas = full[:, o + 2], full[:, o + 3], full[:, o + 4].astype(np.int32)
    from . import lib
    names = ["SCP_SOLVED" if s_ in (0, 1) else f"SCP_FAILED ({lib.CONE_STATUS.get((int(s_) - 2) // 16, '?')})" for s_ in raw]
    from .ptr import SCPBatchSolution
    timing = dict(local.timing) if local is not None else {}
    sol = SCPBatchSolution(names, its, cost, pbm.t, xd, ud, np.ascontiguousarray(p), dev, feas, timing, raw)
    return sol, local
