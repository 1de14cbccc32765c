"""GPU parity: scpb_propagate (final continuous-time trajectory, discretization.jl:515-562) vs the CPU oracle.

Tolerance: fp64, 1e-9 relative to the largest state magnitude of the roll-out (plain RK4 on both sides; the kernel
contracts to FMA, the oracle does not; thousands of steps accumulate rounding differences).  The time arithmetic is
bit-identical, which the starship phase switch (`t <= tau_s`) needs."""
import numpy as np
import pytest

from oracle import orc, problems

pytestmark = pytest.mark.gpu


def _run(handle, name, nb, N, res, seed):
    pb = problems.make_problem(name, N)
    m = pb.orc_model()
    handle.model_set(pb.model_id, pb.par(), pb.nx, pb.nu, pb.np)
    xd, ud, p = problems.test_trajectory(pb, nb, N, seed=seed)
    tc, xc, sec = handle.propagate(orc.t_grid(N), xd, ud, p, res)
    assert xc.shape == (nb, res, pb.nx) and tc[0] == 0.0 and tc[-1] == 1.0 and sec >= 0.0
    for b in range(nb):
        ref = orc.propagate(m, xd[b], ud[b], p[b], res)
        assert np.isfinite(ref).all()
        err = np.abs(xc[b] - ref).max() / max(1.0, np.abs(ref).max())
        assert err <= 1e-9, f"{name} seed {b}: rel err {err:.3e}"
        assert np.array_equal(xc[b][0], xd[b][0])
    return xc


@pytest.mark.parametrize("name,N,Nsub", [("dblint", 9, 10), ("rocket", 8, 15), ("starship", 9, 120),
                                         ("quadrotor", 10, 15), ("freeflyer", 6, 15)])
def test_parity_all_models(handle, name, N, Nsub):
    _run(handle, name, nb=5, N=N, res=2 * Nsub * (N - 1), seed=7)     # res as in scp.jl:231


def test_starship_reference_resolution(handle):
    # starship_flip/tests.jl:33-36 (N = 31, Nsub = 100) -> res = 6000; tau_s = 0.5 is hit exactly at a stage time
    _run(handle, "starship", nb=3, N=31, res=6000, seed=4)


def test_ragged_batch_and_minimal_sizes(handle):
    _run(handle, "quadrotor", nb=33, N=2, res=2, seed=1)      # one step, more seeds than one warp
    _run(handle, "dblint", nb=1, N=3, res=5, seed=2)


def test_freeflyer_quaternion_stays_normalised(handle):
    xc = _run(handle, "freeflyer", nb=2, N=6, res=300, seed=9)
    q = xc[:, 1:, 6:10]
    assert np.abs(np.linalg.norm(q, axis=-1) - 1.0).max() <= 1e-12   # integration action after every step


def test_error_paths(handle, pkg):
    pb = problems.make_problem("dblint", 4)
    handle.model_set(pb.model_id, pb.par(), pb.nx, pb.nu, pb.np)
    xd, ud, p = problems.test_trajectory(pb, 1, 4, seed=0)
    with pytest.raises(pkg.ScpbError):
        handle.propagate(orc.t_grid(4), xd, ud, p, 1)                       # res < 2
    with pytest.raises(pkg.ScpbError):
        handle.propagate(orc.t_grid(4), xd, ud, p, 10, method=pkg.lib.IMPULSE)
