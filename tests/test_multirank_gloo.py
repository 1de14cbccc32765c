"""CPU, world_size 2 and 3 (gloo): the product's multi-GPU entry point scptoolbox.jl_b200/sharded.py -- contiguous seed
blocks of ceil(B / world) per rank, no data-path collective, one all_gather of the per-seed results.
The device loop is replaced by a deterministic stand-in (there is no GPU here); sharding, packing, padding of short
blocks, the collective and the re-assembly are the real code that bench.py and users call."""
import os
import subprocess
import sys
import textwrap

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SCRIPT = textwrap.dedent("""
    import os, sys, numpy as np, torch, torch.distributed as dist
    sys.path.insert(0, %r)
    import __graft_entry__ as g
    pkg = g.load_package()
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    B, N, nx, nu, np_ = 7, 5, 3, 2, 4

    class FakeAlgo:            # stands in for ptr.solve: result is a pure function of the seed's own guess
        calls = []
        @staticmethod
        def solve(pbm, guesses, **kw):
            X, U, P = guesses
            FakeAlgo.calls.append(X.shape[0])
            n = X.shape[0]
            its = (X[:, 0, 0] * 10).astype(np.int32)
            return pkg.ptr.SCPBatchSolution(["SCP_SOLVED"] * n, its, X.sum(axis=(1, 2)), None, 2 * X, 3 * U, P + 1,
                                            np.full(n, 0.5), np.ones(n, dtype=np.int32), {"total": 1.0},
                                            np.zeros(n, dtype=np.int32))

    class Pbm: pass
    pbm = Pbm(); pbm.pars = type("P", (), {"N": N})(); pbm.traj = type("T", (), {"nx": nx, "nu": nu, "np": np_})(); pbm.t = None
    rng = np.random.default_rng(0)           # same full batch on every rank
    X = rng.uniform(0.1, 0.9, (B, N, nx)); U = rng.standard_normal((B, N, nu)); P = rng.standard_normal((B, np_))
    full, local = pkg.sharded.solve_sharded(FakeAlgo, pbm, (X, U, P), dist=dist)
    lo, hi = pkg.sharded.shard_bounds(B, world, rank)
    assert FakeAlgo.calls == ([hi - lo] if hi > lo else []), FakeAlgo.calls      # this rank solved only its own block
    assert np.array_equal(full.xd, 2 * X) and np.array_equal(full.ud, 3 * U) and np.array_equal(full.p, P + 1)
    assert np.array_equal(full.iterations, (X[:, 0, 0] * 10).astype(np.int32)) and np.allclose(full.cost, X.sum(axis=(1, 2)))
    assert full.status == ["SCP_SOLVED"] * B and full.feas.tolist() == [1] * B
    tot = torch.tensor([float(hi - lo)]); dist.all_reduce(tot); assert tot.item() == B    # blocks partition the batch
    dist.destroy_process_group()
    print("rank", rank, "ok")
""") % ROOT


def _run(tmp_path, world, port):
    f = tmp_path / "w.py"
    f.write_text(SCRIPT)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr",
           "127.0.0.1", "--master-port", str(port), str(f)]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert r.returncode == 0, r.stdout + r.stderr
    assert r.stdout.count("ok") == world


def test_two_rank_shard_solve_gather(tmp_path):
    _run(tmp_path, 2, 29731)


def test_three_ranks_with_a_short_last_block(tmp_path):
    _run(tmp_path, 3, 29732)       # 7 seeds -> blocks of 3, 3, 1


def test_shard_bounds_cover_the_batch(pkg):
    for B in (1, 7, 256, 257):
        for world in (1, 2, 3, 8):
            blocks = [pkg.sharded.shard_bounds(B, world, r) for r in range(world)]
            assert blocks[0][0] == 0 and blocks[-1][1] == B
            assert all(a[1] == b[0] for a, b in zip(blocks, blocks[1:]))
            assert max(hi - lo for lo, hi in blocks) == -(-B // world)
    assert pkg.sharded.shard_bounds(256, 8, 3) == (96, 128)      # 32 seeds per GPU on 8 GPUs
