"""CPU, world_size 2 (gloo): the multi-GPU plumbing of bench.py -- disjoint per-rank seed blocks (no data-path
collective), MAX-over-ranks timing, SUM of iteration counts, final all_gather of the per-seed results."""
import os
import subprocess
import sys
import textwrap

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SCRIPT = textwrap.dedent("""
    import os, sys, numpy as np, torch, torch.distributed as dist
    sys.path.insert(0, %r)
    import bench
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    base = (np.ones((5, 3)), np.ones((5, 2)), np.ones(4))
    X, U, P = bench.make_seeds(base, np.ones(3), np.ones(2), 6, rank)
    # every rank owns different seeds
    h = torch.tensor([float(X.sum())]); allh = [torch.zeros(1) for _ in range(world)]
    dist.all_gather(allh, h)
    assert abs(allh[0].item() - allh[1].item()) > 1e-9
    t = torch.tensor([1.0 + rank, 2.0 - rank], dtype=torch.float64); dist.all_reduce(t, op=dist.ReduceOp.MAX)
    assert t.tolist() == [2.0, 2.0]
    cnt = torch.tensor([10.0 * (rank + 1)], dtype=torch.float64); dist.all_reduce(cnt, op=dist.ReduceOp.SUM)
    assert cnt.item() == 30.0
    res = torch.from_numpy(np.concatenate([X.reshape(6, -1), U.reshape(6, -1), P], axis=1))
    out = [torch.empty_like(res) for _ in range(world)]
    dist.all_gather(out, res)
    assert torch.equal(out[rank], res) and not torch.equal(out[0], out[1])
    dist.destroy_process_group()
    print("rank", rank, "ok")
""") % ROOT


def test_two_rank_sharding_and_gather(tmp_path):
    f = tmp_path / "w.py"
    f.write_text(SCRIPT)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr",
           "127.0.0.1", "--master-port", "29731", str(f)]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert r.returncode == 0, r.stdout + r.stderr
    assert r.stdout.count("ok") == 2
