"""CPU: the product's symbolic PTR template (scptoolbox.jl_b200/ptr.py + parser.py) evaluated at a reference
must reproduce the oracle's numeric subproblem (oracle/ptr.py, restating ptr.jl / scp.jl) exactly:
    [Avals; Gvals; c; b; h] = W @ src   vs   oracle ConeProgram.compile()."""
import numpy as np
import pytest
import scipy.sparse as sp

from oracle import orc, problems, ptr as optr


def _sources(sm, pbo, P, ref):
    """Fill the source vector on the host from oracle quantities (what the device kernels compute)."""
    N, nx, nu, np_, ns, nf = sm.N, sm.nx, sm.nu, sm.np, sm.ns, sm.nf
    ng = getattr(sm, "ng", np_)
    gcols = getattr(pbo, "gcols", None) or (lambda k: list(range(np_)))
    src = np.zeros(sm.nsrc); src[0] = 1.0
    d = ref.dyn
    for k in range(N - 1):
        src[sm.oA + k * nx * nx: sm.oA + (k + 1) * nx * nx] = d.A[k].flatten(order="F")
        src[sm.oBm + k * nx * nu: sm.oBm + (k + 1) * nx * nu] = d.Bm[k].flatten(order="F")
        src[sm.oBp + k * nx * nu: sm.oBp + (k + 1) * nx * nu] = d.Bp[k].flatten(order="F")
        src[sm.oF + k * nx * nf: sm.oF + (k + 1) * nx * nf] = d.F[k][:, :nf].flatten(order="F")
        src[sm.or_ + k * nx: sm.or_ + (k + 1) * nx] = d.r[k]
        src[sm.oE + k * nx * nx: sm.oE + (k + 1) * nx * nx] = d.E[k].flatten(order="F")
    t = P.t
    for k in range(N):
        if ns:
            a = (t[k], k + 1, ref.xd[k], ref.ud[k], ref.p)
            Cm, Dm, Gm, s = pbo.C(*a), pbo.D(*a), pbo.G(*a), pbo.s(*a)
            src[sm.oC + k * ns * nx: sm.oC + (k + 1) * ns * nx] = Cm.flatten()
            src[sm.oD + k * ns * nu: sm.oD + (k + 1) * ns * nu] = Dm.flatten()
            src[sm.oG + k * ns * ng: sm.oG + (k + 1) * ns * ng] = Gm[:, gcols(k)].flatten()
            src[sm.ors + k * ns: sm.ors + (k + 1) * ns] = s - Cm @ ref.xd[k] - Dm @ ref.ud[k] - Gm @ ref.p
        src[sm.oxh + k * nx: sm.oxh + (k + 1) * nx] = (ref.xd[k] - P.scale.cx) * P.scale.iSx
        src[sm.ouh + k * nu: sm.ouh + (k + 1) * nu] = (ref.ud[k] - P.scale.cu) * P.scale.iSu
    src[sm.oph: sm.oph + np_] = (ref.p - P.scale.cp) * P.scale.iSp
    return src


class _NoGpu:
    """stand-in for the device objects so that the template can be built on a CPU-only box"""


@pytest.mark.parametrize("N", [9, 14])
def test_template_matches_oracle_subproblem(pkg, N, monkeypatch):
    # oracle side
    pbo = problems.StarshipProblem(N)
    rng = np.random.default_rng(N)
    xd, ud, p = problems.test_trajectory(pbo, 1, N, seed=N)
    pbo.hs = 123.4
    opars = optr.Parameters(N=N, Nsub=60, iter_max=5, wvc=1e3, wtr=0.1, eps_abs=1e-5, eps_rel=1e-4, feas_tol=5e-3)
    P = optr.PTR(pbo, opars)
    ref = P.make_solution(xd[0], ud[0], p[0])
    prg, _ = P.build(ref)
    ocp = prg.compile()
    # product side: build the template without touching the GPU
    ex = pkg.examples.starship
    mdl = ex.StarshipProblem(); mdl.hs = 123.4
    traj = pkg.problem.TrajectoryProblem(mdl)
    ex.define_problem(traj, "ptr", handle=None)
    pars = pkg.ptr.Parameters(N=N, Nsub=60, iter_max=5, disc_method=pkg.ptr.FOH, wvc=1e3, wtr=0.1, eps_abs=1e-5,
                              eps_rel=1e-4, feas_tol=5e-3, q_tr=np.inf, q_exit=np.inf)

    class FakeHandle:
        def model_set(self, *a): pass
    monkeypatch.setattr(pkg.lib, "ConeProblem", lambda *a, **k: type("C", (), {"c": None, "close": lambda s: None})())
    fake = FakeHandle(); fake.lib = type("L", (), {"scpb_ptr_setup": staticmethod(lambda *a: 0)})(); fake.h = None
    fake._check = lambda rc, what: None
    pbm = pkg.ptr.SCPProblem(pars, traj, fake, l1_block=0)   # the reference's exact (NormOneBridge) program
    cp, sm = pbm.cp, pbm.sm
    src = _sources(sm, pbo, P, ref)
    vals = pbm.W @ src
    n, p_, m = cp["n"], cp["p"], cp["m"]
    assert (n, p_, m, cp["l"]) == (ocp["c"].size, ocp["A"].shape[0], ocp["G"].shape[0], ocp["l"])
    A = sp.csr_matrix((vals[:cp["nnzA"]], cp["A"].indices, cp["A"].indptr), shape=(p_, n))
    G = sp.csr_matrix((vals[cp["nnzA"]:cp["nnzA"] + cp["nnzG"]], cp["G"].indices, cp["G"].indptr), shape=(m, n))
    tol = 1e-12
    assert abs(A - ocp["A"]).max() <= tol * max(1.0, abs(ocp["A"]).max())
    assert abs(G - ocp["G"]).max() <= tol * max(1.0, abs(ocp["G"]).max())
    c = vals[cp["off_c"]:cp["off_c"] + n]; b = vals[cp["off_b"]:cp["off_b"] + p_]; h = vals[cp["off_h"]:cp["off_h"] + m]
    assert np.abs(c - ocp["c"]).max() <= tol * max(1.0, np.abs(ocp["c"]).max())
    assert np.abs(b - ocp["b"]).max() <= 1e-11 * max(1.0, np.abs(ocp["b"]).max())
    assert np.abs(h - ocp["h"]).max() <= 1e-11 * max(1.0, np.abs(ocp["h"]).max())
    assert abs(vals[-1] - ocp["c0"]) <= 1e-12
    # stage ordering covers every KKT node exactly once
    perm = pbm.perm
    assert sorted(perm.tolist()) == list(range(n + p_))


def test_rocket_template_matches_oracle_subproblem(pkg, monkeypatch):
    """Same check for the SOC-constrained PTR case (rocket landing, BASELINE config C2): two SOC(4) cones per node."""
    N = 10
    pbo = problems.RocketProblem(N)
    xd, ud, p = problems.test_trajectory(pbo, 1, N, seed=3)
    opars = optr.Parameters(N=N, Nsub=15, iter_max=5, wvc=1e3, wtr=0.1, eps_abs=1e-5, eps_rel=1e-4, feas_tol=1e-3)
    P = optr.PTR(pbo, opars)
    ref = P.make_solution(xd[0], ud[0], p[0])
    prg, _ = P.build(ref)
    ocp = prg.compile()
    ex = pkg.examples.rocket_landing
    mdl = ex.RocketProblem()
    traj = pkg.problem.TrajectoryProblem(mdl)
    ex.define_problem(traj, "ptr", handle=None)
    pars = pkg.ptr.Parameters(N=N, Nsub=15, iter_max=5, disc_method=pkg.ptr.FOH, wvc=1e3, wtr=0.1, eps_abs=1e-5,
                              eps_rel=1e-4, feas_tol=1e-3, q_tr=np.inf, q_exit=np.inf)

    class FakeHandle:
        def model_set(self, *a): pass
    monkeypatch.setattr(pkg.lib, "ConeProblem", lambda *a, **k: type("C", (), {"c": None, "close": lambda s: None})())
    fake = FakeHandle(); fake.lib = type("L", (), {"scpb_ptr_setup": staticmethod(lambda *a: 0)})(); fake.h = None
    fake._check = lambda rc, what: None
    pbm = pkg.ptr.SCPProblem(pars, traj, fake, l1_block=0)
    cp, sm = pbm.cp, pbm.sm
    vals = pbm.W @ _sources(sm, pbo, P, ref)
    n, p_, m = cp["n"], cp["p"], cp["m"]
    assert (n, p_, m, cp["l"]) == (ocp["c"].size, ocp["A"].shape[0], ocp["G"].shape[0], ocp["l"])
    assert list(cp["soc_dims"]) == list(ocp["q"]) == [4] * (2 * N)
    A = sp.csr_matrix((vals[:cp["nnzA"]], cp["A"].indices, cp["A"].indptr), shape=(p_, n))
    G = sp.csr_matrix((vals[cp["nnzA"]:cp["nnzA"] + cp["nnzG"]], cp["G"].indices, cp["G"].indptr), shape=(m, n))
    tol = 1e-12
    assert abs(A - ocp["A"]).max() <= tol * max(1.0, abs(ocp["A"]).max())
    assert abs(G - ocp["G"]).max() <= tol * max(1.0, abs(ocp["G"]).max())
    c = vals[cp["off_c"]:cp["off_c"] + n]; b = vals[cp["off_b"]:cp["off_b"] + p_]; h = vals[cp["off_h"]:cp["off_h"] + m]
    assert np.abs(c - ocp["c"]).max() <= tol * max(1.0, np.abs(ocp["c"]).max())
    assert np.abs(b - ocp["b"]).max() <= 1e-11 * max(1.0, np.abs(ocp["b"]).max())
    assert np.abs(h - ocp["h"]).max() <= 1e-11 * max(1.0, np.abs(ocp["h"]).max())
    assert abs(vals[-1] - ocp["c0"]) <= 1e-12


@pytest.mark.parametrize("q_tr", [np.inf, 2, 4])
def test_scvx_template_matches_oracle_subproblem(pkg, monkeypatch, q_tr):
    """SCvx flavour of the template (trust-region radius as a device source, lambda penalty, no eta variables) against
    the oracle's numeric SCvx subproblem (oracle/scvx.py restating scvx.jl:225-303, 578-701, 804-901), for the LINF, SOC and
    squared-two-norm (GEOM, scvx.jl:646-662) trust regions."""
    from oracle import scvx as oscvx
    N = 9
    pbo = problems.StarshipProblem(N)
    xd, ud, p = problems.test_trajectory(pbo, 1, N, seed=5)
    pbo.hs = 77.0
    kw = dict(lam=5e2, rho_0=0.0, rho_1=0.1, rho_2=0.7, beta_sh=2.0, beta_gr=2.0, eta_init=1.0, eta_lb=1e-8, eta_ub=10.0,
              eps_abs=1e-5, eps_rel=1e-4, feas_tol=5e-3)
    S = oscvx.SCvx(pbo, oscvx.Parameters(N=N, Nsub=60, iter_max=5, q_tr=q_tr, **kw))
    ref = S.make_solution(xd[0], ud[0], p[0])
    eta = 0.37
    prg, _ = S.build(ref, eta)
    ocp = prg.compile()
    ex = pkg.examples.starship
    mdl = ex.StarshipProblem(); mdl.hs = 77.0
    traj = pkg.problem.TrajectoryProblem(mdl)
    ex.define_problem(traj, "scvx", handle=None)
    pars = pkg.scvx.Parameters(N=N, Nsub=60, iter_max=5, disc_method=pkg.ptr.FOH, q_tr=q_tr, q_exit=np.inf, **kw)

    class FakeHandle:
        def model_set(self, *a): pass
    monkeypatch.setattr(pkg.lib, "ConeProblem", lambda *a, **k: type("C", (), {"c": None, "close": lambda s: None})())
    fake = FakeHandle(); fake.lib = type("L", (), {"scpb_ptr_setup": staticmethod(lambda *a: 0)})(); fake.h = None
    fake._check = lambda rc, what: None
    pbm = pkg.ptr.SCPProblem(pars, traj, fake, l1_block=0, algo="scvx")
    cp, sm = pbm.cp, pbm.sm
    src = _sources(sm, pbo, S, ref)
    src[sm.oeta] = eta
    vals = pbm.W @ src
    n, p_, m = cp["n"], cp["p"], cp["m"]
    assert (n, p_, m, cp["l"]) == (ocp["c"].size, ocp["A"].shape[0], ocp["G"].shape[0], ocp["l"])
    A = sp.csr_matrix((vals[:cp["nnzA"]], cp["A"].indices, cp["A"].indptr), shape=(p_, n))
    G = sp.csr_matrix((vals[cp["nnzA"]:cp["nnzA"] + cp["nnzG"]], cp["G"].indices, cp["G"].indptr), shape=(m, n))
    tol = 1e-12
    assert abs(A - ocp["A"]).max() <= tol * max(1.0, abs(ocp["A"]).max())
    assert abs(G - ocp["G"]).max() <= tol * max(1.0, abs(ocp["G"]).max())
    c = vals[cp["off_c"]:cp["off_c"] + n]; b = vals[cp["off_b"]:cp["off_b"] + p_]; h = vals[cp["off_h"]:cp["off_h"] + m]
    assert np.abs(c - ocp["c"]).max() <= tol * max(1.0, np.abs(ocp["c"]).max())
    assert np.abs(b - ocp["b"]).max() <= 1e-11 * max(1.0, np.abs(ocp["b"]).max())
    assert np.abs(h - ocp["h"]).max() <= 1e-11 * max(1.0, np.abs(ocp["h"]).max())
    assert abs(vals[-1] - ocp["c0"]) <= 1e-12
    # Q rows (original cost, boundary conditions) evaluated at the scaled reference = the oracle's numeric values
    rp, ci, v, c0 = pkg.scvx._rows_matrix([pkg.parser.Expr.lift(pbm.J_orig)] + pbm.g_ic + pbm.g_tc, n)
    z = np.zeros(n)
    sc = pbm.scale
    ox, ou, op = pbm.template.blocks["x"][0], pbm.template.blocks["u"][0], pbm.template.blocks["p"][0]
    z[ox:ox + N * 8] = ((ref.xd - sc.cx) / sc.Sx).ravel()
    z[ou:ou + N * 3] = ((ref.ud - sc.cu) / sc.Su).ravel()
    z[op:op + 10] = (ref.p - sc.cp) / sc.Sp
    Q = sp.csr_matrix((v, ci, rp), shape=(len(c0), n))
    q = Q @ z + c0
    want = np.concatenate([[S.original_cost(ref)], pbo.gic(ref.xd[0], ref.p), pbo.gtc(ref.xd[-1], ref.p)])
    assert np.abs(q - want).max() <= 1e-10 * max(1.0, np.abs(want).max())


@pytest.mark.parametrize("q_tr", [1, 2, 4])
def test_template_other_trust_region_norms(pkg, monkeypatch, q_tr):
    """q_tr in {1, 2, 4} (ptr.jl:582: L1 / SOC trust-region cones instead of LINF; 4 = squared two-norm through the GEOM
    cone, ptr.jl:604-630): template == oracle subproblem."""
    N = 7
    pbo = problems.StarshipProblem(N)
    xd, ud, p = problems.test_trajectory(pbo, 1, N, seed=q_tr)
    pbo.hs = 50.0
    opars = optr.Parameters(N=N, Nsub=40, iter_max=5, wvc=1e3, wtr=0.1, eps_abs=1e-5, eps_rel=1e-4, feas_tol=5e-3, q_tr=q_tr)
    P = optr.PTR(pbo, opars)
    ref = P.make_solution(xd[0], ud[0], p[0])
    prg, _ = P.build(ref)
    ocp = prg.compile()
    ex = pkg.examples.starship
    mdl = ex.StarshipProblem(); mdl.hs = 50.0
    traj = pkg.problem.TrajectoryProblem(mdl)
    ex.define_problem(traj, "ptr", handle=None)
    pars = pkg.ptr.Parameters(N=N, Nsub=40, iter_max=5, disc_method=pkg.ptr.FOH, wvc=1e3, wtr=0.1, eps_abs=1e-5,
                              eps_rel=1e-4, feas_tol=5e-3, q_tr=q_tr, q_exit=np.inf)

    class FakeHandle:
        def model_set(self, *a): pass
    monkeypatch.setattr(pkg.lib, "ConeProblem", lambda *a, **k: type("C", (), {"c": None, "close": lambda s: None})())
    fake = FakeHandle(); fake.lib = type("L", (), {"scpb_ptr_setup": staticmethod(lambda *a: 0)})(); fake.h = None
    fake._check = lambda rc, what: None
    pbm = pkg.ptr.SCPProblem(pars, traj, fake, l1_block=0)
    cp, sm = pbm.cp, pbm.sm
    vals = pbm.W @ _sources(sm, pbo, P, ref)
    n, p_, m = cp["n"], cp["p"], cp["m"]
    assert (n, p_, m, cp["l"]) == (ocp["c"].size, ocp["A"].shape[0], ocp["G"].shape[0], ocp["l"])
    assert list(cp["soc_dims"]) == list(ocp["q"])
    if q_tr == 2:
        assert sorted(set(cp["soc_dims"])) == [4, 9, 11]      # input, state and parameter trust regions
    if q_tr == 4:
        assert sorted(set(cp["soc_dims"])) == [2, 3, 4, 9, 11]   # ... plus |d_lq| <= w and the lowered GEOM cones
    A = sp.csr_matrix((vals[:cp["nnzA"]], cp["A"].indices, cp["A"].indptr), shape=(p_, n))
    G = sp.csr_matrix((vals[cp["nnzA"]:cp["nnzA"] + cp["nnzG"]], cp["G"].indices, cp["G"].indptr), shape=(m, n))
    tol = 1e-12
    assert abs(A - ocp["A"]).max() <= tol * max(1.0, abs(ocp["A"]).max())
    assert abs(G - ocp["G"]).max() <= tol * max(1.0, abs(ocp["G"]).max())
    c = vals[cp["off_c"]:cp["off_c"] + n]; b = vals[cp["off_b"]:cp["off_b"] + p_]; h = vals[cp["off_h"]:cp["off_h"] + m]
    assert np.abs(c - ocp["c"]).max() <= tol * max(1.0, np.abs(ocp["c"]).max())
    assert np.abs(b - ocp["b"]).max() <= 1e-11 * max(1.0, np.abs(ocp["b"]).max())
    assert np.abs(h - ocp["h"]).max() <= 1e-11 * max(1.0, np.abs(ocp["h"]).max())


def test_l1_block_lowering_is_an_equivalent_program(pkg, monkeypatch):
    """The product lowers wide L1 cones through partial sums (l1_block = 4, keeps the KKT factor sparse); the reference's
    NormOneBridge form is l1_block = 0.  Both must be the same optimisation problem: equal optimal value and equal
    (x, u, p) minimiser on the oracle's data (HiGHS on both compiled programs)."""
    from oracle import conic
    N = 8
    pbo = problems.StarshipProblem(N)
    xd, ud, p = problems.test_trajectory(pbo, 1, N, seed=2)
    pbo.hs = 60.0
    opars = optr.Parameters(N=N, Nsub=40, iter_max=5, wvc=1e3, wtr=0.1, eps_abs=1e-5, eps_rel=1e-4, feas_tol=5e-3)
    P = optr.PTR(pbo, opars)
    ref = P.make_solution(xd[0], ud[0], p[0])
    ex = pkg.examples.starship
    mdl = ex.StarshipProblem(); mdl.hs = 60.0
    traj = pkg.problem.TrajectoryProblem(mdl)
    ex.define_problem(traj, "ptr", handle=None)
    pars = pkg.ptr.Parameters(N=N, Nsub=40, iter_max=5, disc_method=pkg.ptr.FOH, wvc=1e3, wtr=0.1, eps_abs=1e-5,
                              eps_rel=1e-4, feas_tol=5e-3, q_tr=np.inf, q_exit=np.inf)

    class FakeHandle:
        def model_set(self, *a): pass
    monkeypatch.setattr(pkg.lib, "ConeProblem", lambda *a, **k: type("C", (), {"c": None, "close": lambda s: None})())
    fake = FakeHandle(); fake.lib = type("L", (), {"scpb_ptr_setup": staticmethod(lambda *a: 0)})(); fake.h = None
    fake._check = lambda rc, what: None
    res = {}
    for blk in (0, 4):
        pbm = pkg.ptr.SCPProblem(pars, traj, fake, l1_block=blk)
        cp, sm = pbm.cp, pbm.sm
        vals = pbm.W @ _sources(sm, pbo, P, ref)
        n, p_, m = cp["n"], cp["p"], cp["m"]
        A = sp.csr_matrix((vals[:cp["nnzA"]], cp["A"].indices, cp["A"].indptr), shape=(p_, n))
        G = sp.csr_matrix((vals[cp["nnzA"]:cp["nnzA"] + cp["nnzG"]], cp["G"].indices, cp["G"].indptr), shape=(m, n))
        prog = dict(c=vals[cp["off_c"]:cp["off_c"] + n], c0=vals[-1], A=A, b=vals[cp["off_b"]:cp["off_b"] + p_], G=G,
                    h=vals[cp["off_h"]:cp["off_h"] + m], l=cp["l"], q=[])
        out = conic.solve_highs(prog, tol=1e-9)
        assert out["status"] == "OPTIMAL"
        nxu = N * (8 + 3) + 10                      # x, u, p blocks come first in both programs
        res[blk] = (out["obj"], out["z"][:nxu], n)
    assert res[4][2] > res[0][2]                    # the lowering adds partial-sum variables ...
    assert abs(res[0][0] - res[4][0]) <= 1e-8 * max(1.0, abs(res[0][0]))          # ... and changes nothing else
    d = np.abs(res[0][1] - res[4][1])
    assert np.median(d) <= 1e-7                     # (flat directions may differ between two LP vertices)


def test_freeflyer_template_matches_oracle_subproblem(pkg, monkeypatch):
    """BASELINE config C5 (free-flyer PTR): SOC(4) velocity / rate / thrust / torque cones, LINF room SDF cones, the
    log-sum-exp SDF row with its packed ds/dp columns (np = 1 + 6N), the quadratic running cost lowered to one rotated
    cone per node.  The bounding boxes come from the oracle's compute_bbox (the product computes its own on the GPU,
    tests/test_ptr_gpu.py compares the two)."""
    N = 6
    pbo = problems.FreeFlyerProblem(N)
    pbo.gcols = lambda k: [pbo.idd(i, k) for i in range(pbo.n_iss)]
    xd, ud, p = pbo.guess(N)
    rng = np.random.default_rng(2)
    xd = xd + 1e-3 * rng.standard_normal(xd.shape); ud = ud + 1e-4 * rng.standard_normal(ud.shape)
    opars = optr.Parameters(N=N, Nsub=15, iter_max=5, wvc=1e3, wtr=0.1, eps_abs=1e-5, eps_rel=1e-4, feas_tol=1e-3)
    P = optr.PTR(pbo, opars)
    ref = P.make_solution(xd, ud, p)
    prg, _ = P.build(ref)
    ocp = prg.compile()
    ex = pkg.examples.freeflyer
    mdl = ex.FreeFlyerProblem(N)
    traj = pkg.problem.TrajectoryProblem(mdl)
    ex.define_problem(traj, "ptr", handle=None)
    xb, ub, pb_ = optr.compute_bbox(pbo, N)           # advise every variable: no GPU on this box
    for i, r in enumerate(xb): pkg.problem.problem_advise_scale(traj, "state", i, r)
    for i, r in enumerate(ub): pkg.problem.problem_advise_scale(traj, "input", i, r)
    pars = pkg.ptr.Parameters(N=N, Nsub=15, iter_max=5, disc_method=pkg.ptr.FOH, wvc=1e3, wtr=0.1, eps_abs=1e-5,
                              eps_rel=1e-4, feas_tol=1e-3, q_tr=np.inf, q_exit=np.inf)

    class FakeHandle:
        def model_set(self, *a): pass
    monkeypatch.setattr(pkg.lib, "ConeProblem", lambda *a, **k: type("C", (), {"c": None, "close": lambda s: None})())
    fake = FakeHandle(); fake.lib = type("L", (), {"scpb_ptr_setup": staticmethod(lambda *a: 0)})(); fake.h = None
    fake._check = lambda rc, what: None
    pbm = pkg.ptr.SCPProblem(pars, traj, fake, l1_block=0)
    cp, sm = pbm.cp, pbm.sm
    assert sm.ng == 6 and pbm.desc.ng == 6
    # same guess on both sides
    xg, ug, pg = traj.guess(N)
    x0, u0, p0 = pbo.guess(N)
    assert np.abs(xg - x0).max() < 1e-12 and np.abs(pg - p0).max() < 1e-12
    assert np.abs(pbm.scale.Sx - P.scale.Sx).max() < 1e-9 and np.abs(pbm.scale.Su - P.scale.Su).max() < 1e-9
    vals = pbm.W @ _sources(sm, pbo, P, ref)
    n, p_, m = cp["n"], cp["p"], cp["m"]
    assert (n, p_, m, cp["l"]) == (ocp["c"].size, ocp["A"].shape[0], ocp["G"].shape[0], ocp["l"])
    assert sorted(cp["soc_dims"]) == sorted(ocp["q"])
    # variables and rows are emitted in a different order on the two sides (epigraph variables, cone order): compare
    # the programs through their optimal values instead of entry by entry
    from oracle import conic
    A = sp.csr_matrix((vals[:cp["nnzA"]], cp["A"].indices, cp["A"].indptr), shape=(p_, n))
    G = sp.csr_matrix((vals[cp["nnzA"]:cp["nnzA"] + cp["nnzG"]], cp["G"].indices, cp["G"].indptr), shape=(m, n))
    mine = dict(c=vals[cp["off_c"]:cp["off_c"] + n], c0=vals[-1], A=A, b=vals[cp["off_b"]:cp["off_b"] + p_], G=G,
                h=vals[cp["off_h"]:cp["off_h"] + m], l=cp["l"], q=list(cp["soc_dims"]))
    r1 = conic.solve_ipm(mine, tol=1e-9)
    r2 = conic.solve_ipm(ocp, tol=1e-9)
    assert r1["status"] == r2["status"] == "OPTIMAL"
    assert abs(r1["obj"] - r2["obj"]) <= 1e-7 * max(1.0, abs(r2["obj"]))
    bx = pbm.template.blocks
    ox, (nx_, _) = bx["x"][0], bx["x"][1]
    oxo = prg.blocks["x"][0]
    assert np.abs(r1["z"][ox:ox + nx_ * N] - r2["z"][oxo:oxo + nx_ * N]).max() <= 1e-5


def test_gusto_template_matches_oracle_subproblem(pkg, monkeypatch):
    """BASELINE config C4 (quadrotor GuSTO): the GuSTO flavour of the template (un-relaxed dynamics and boundary
    conditions, quadratic soft penalties through rotated cones with the per-seed sources sqrt(lambda) and eta) filled with
    oracle quantities is the oracle's GuSTO subproblem: same sizes, same optimal value, same trajectory.  Also checks the
    rows handed to scpb_gusto_attach: J = affine + weighted squares reproduces the oracle's original_cost, the last row is
    L_tr."""
    from oracle import conic, gusto as ogusto
    N = 10
    pbo = problems.QuadrotorProblem(N)
    xd, ud, p = pbo.guess(N)
    rng = np.random.default_rng(4)
    xd = xd + 1e-2 * rng.standard_normal(xd.shape); ud = ud + 1e-2 * rng.standard_normal(ud.shape)
    kw = dict(lam_init=1e4, lam_max=1e9, rho_0=0.1, rho_1=0.9, beta_sh=2.0, beta_gr=2.0, gamma_fail=5.0, eta_init=10.0,
              eta_lb=1e-3, eta_ub=10.0, mu=0.8, iter_mu=6, eps_abs=0.0, eps_rel=0.0, feas_tol=1e-3)
    P = ogusto.GuSTO(pbo, ogusto.Parameters(N=N, Nsub=15, iter_max=15, **kw))
    ref = P.make_solution(xd, ud, p)
    lam, eta = 5e4, 0.7
    prg, hnd = P.build(ref, lam, eta)
    ocp = prg.compile()
    ex = pkg.examples.quadrotor
    mdl = ex.QuadrotorProblem()
    traj = pkg.problem.TrajectoryProblem(mdl)
    ex.define_problem(traj, "gusto", handle=None)
    xb, ub, pb_ = optr.compute_bbox(pbo, N)           # advise every variable: no GPU on this box
    for i, r in enumerate(xb): pkg.problem.problem_advise_scale(traj, "state", i, r)
    for i, r in enumerate(ub): pkg.problem.problem_advise_scale(traj, "input", i, r)
    pars = pkg.gusto.Parameters(N, 15, 15, pkg.ptr.FOH, kw["lam_init"], kw["lam_max"], kw["rho_0"], kw["rho_1"], kw["beta_sh"],
                                kw["beta_gr"], kw["gamma_fail"], kw["eta_init"], kw["eta_lb"], kw["eta_ub"], kw["mu"],
                                kw["iter_mu"], 0.0, 0.0, 1e-3, "quad", 100.0, np.inf, np.inf)

    class FakeHandle:
        def model_set(self, *a): pass
    monkeypatch.setattr(pkg.lib, "ConeProblem", lambda *a, **k: type("C", (), {"c": None, "close": lambda s: None})())
    fake = FakeHandle()
    fake.lib = type("L", (), {"scpb_ptr_setup": staticmethod(lambda *a: 0), "scpb_gusto_attach": staticmethod(lambda *a: 0)})()
    fake.h = None
    fake._check = lambda rc, what: None
    pbm = pkg.gusto.GuSTOProblem(pars, traj, fake, l1_block=0)
    cp, sm = pbm.cp, pbm.sm
    assert np.abs(pbm.scale.Sx - P.scale.Sx).max() < 1e-12 and np.abs(pbm.scale.Su - P.scale.Su).max() < 1e-9
    src = _sources(sm, pbo, P, ref)
    src[sm.oeta] = eta; src[sm.olam] = lam
    vals = pbm.W @ src
    n, p_, m = cp["n"], cp["p"], cp["m"]
    assert (n, p_, m, cp["l"]) == (ocp["c"].size, ocp["A"].shape[0], ocp["G"].shape[0], ocp["l"])
    assert sorted(cp["soc_dims"]) == sorted(ocp["q"])
    A = sp.csr_matrix((vals[:cp["nnzA"]], cp["A"].indices, cp["A"].indptr), shape=(p_, n))
    G = sp.csr_matrix((vals[cp["nnzA"]:cp["nnzA"] + cp["nnzG"]], cp["G"].indices, cp["G"].indptr), shape=(m, n))
    mine = dict(c=vals[cp["off_c"]:cp["off_c"] + n], c0=vals[-1], A=A, b=vals[cp["off_b"]:cp["off_b"] + p_], G=G,
                h=vals[cp["off_h"]:cp["off_h"] + m], l=cp["l"], q=list(cp["soc_dims"]))
    r1 = conic.solve_ipm(mine, tol=1e-9)
    r2 = conic.solve_ipm(ocp, tol=1e-9)
    print("gusto template:", r1["status"], r2["status"], r1["obj"], r2["obj"])
    assert r1["status"] in ("OPTIMAL", "ALMOST_OPTIMAL") and r2["status"] in ("OPTIMAL", "ALMOST_OPTIMAL")
    assert abs(r1["obj"] - r2["obj"]) <= 1e-6 * max(1.0, abs(r2["obj"]))     # two runs of a tol-1e-9 solver, |obj| ~ 1e6
    bx = pbm.template.blocks
    for name, cnt in (("x", 6 * N), ("u", 4 * N), ("p", 1)):
        o1, o2 = bx[name][0], prg.blocks[name][0]
        assert np.abs(r1["z"][o1:o1 + cnt] - r2["z"][o2:o2 + cnt]).max() <= 1e-5, name
    # the rows of scpb_gusto_attach at the oracle's solution
    z2 = r2["z"]
    val = np.vectorize(lambda e: e.value(z2), otypes=[float])
    xs, us, ps = val(hnd["x"]).T, val(hnd["u"]).T, val(hnd["p"])
    sc = pbm.scale
    zq = np.zeros(n)
    zq[bx["x"][0]:bx["x"][0] + 6 * N] = ((xs - sc.cx) / sc.Sx).ravel()
    zq[bx["u"][0]:bx["u"][0] + 4 * N] = ((us - sc.cu) / sc.Su).ravel()
    zq[bx["p"][0]] = (ps[0] - sc.cp[0]) / sc.Sp[0]
    rp, ci, v, c0 = pbm.Q
    nsq = pbm.gdesc.nsq
    rows = np.array([c0[r] + v[rp[r]:rp[r + 1]] @ zq[ci[rp[r]:rp[r + 1]]] for r in range(nsq + 1)])
    J = rows[0] + (pbm.Q_w[:nsq] * rows[1:] ** 2).sum()
    assert abs(J - P.original_cost(xs, us, ps)) <= 1e-12 * max(1.0, abs(J))
    r = nsq + 1
    Ltr = lam * (c0[r] + v[rp[r]:rp[r + 1]] @ r1["z"][ci[rp[r]:rp[r + 1]]])
    assert abs(Ltr - conic.Aff.lift(hnd["L_tr"]).value(z2)) <= 1e-7 * max(1.0, abs(Ltr))
