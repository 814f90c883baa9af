"""GPU parity on the meshes BASELINE.json's configs name, against the REFERENCE ITSELF (oracle/_ref, see tests/refcases.py):

  * config 2 -- short_tests/ethier: the 27-element box on [-1,1]^3 with 'v  ' on every side;
  * config 5 -- examples/turbChannel: periodic in x and z, walls in y, vertices stretched by the case's usrdat; a 4^3-element
    cut with full golden fields, and the complete 16 x 12 x 8 = 1536-element mesh with sampled golden fields.

Tolerances are the north star's: identical CG / GMRES iteration counts, fields <= 1e-10 relative (a solve run to
convergence is compared at the accuracy it was asked for, as in tests/test_gpu_golden.py).  The file sorts after the other
GPU suites on purpose: these are the largest cases.
"""
import os

import numpy as np
import pytest

import refcases
from oracle import hsmg

pytestmark = pytest.mark.gpu

TOL_FIELD = 1e-10
TOL_CONVERGED = 1e-7

G = refcases.load_golden()


def relmax(a, b):
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-300)


@pytest.fixture()
def nek():
    from nek5000_b200 import nek as N
    N.finalize()
    N.init(0, 8, 3)
    yield N
    N.finalize()


def _register(nek, case, geo, bm1, binv, volume, zg, wg, D):
    E = case.nel
    nek.set_nel(E, E)
    nek.set_gll(zg, wg)
    nek.set_dxyz(D, np.ascontiguousarray(D.T))
    nek.set_geom(*geo, bm1)
    nek.set_ifdfrm(None)
    h, _ = nek.setupds(8, E, case.vertex)
    nek.set_ifield(1)
    nek.set_field_handle(1, h)
    nek.set_step_info(1, float(volume))
    nek.set_binv(binv)


@pytest.mark.parametrize("name", ["ethier", "channel"])
def test_config_mesh_velocity_and_pressure_solves_against_the_reference(nek, name):
    g = G[name]
    case = refcases.channel_case() if name == "channel" else refcases.ethier_case()
    fbc = refcases.channel_fbc(case) if name == "channel" else hsmg.box_fbc(case, (2,) * 6)
    E, n = case.nel, case.n
    _register(nek, case, [g[f"g{i}m1"] for i in range(1, 7)], g["bm1"], g["binvm1"], g["volvm1"][0], g["zgm1"], g["wxm1"], g["dxm1"])
    # velocity: hmholtz('VELX') with the constants of a time step (h1 = viscosity, h2 = bd/dt)
    nek.set_param(22, 0.0)
    x, rhs = np.zeros(n), g["vel_rhs"].copy()
    it = nek.hmholtz("VELX", x, rhs, np.full(n, g["vel_h1"][0]), np.full(n, g["vel_h2"][0]), g["v1mask"], g["vmult"], 1, 1e-9, 200, 1)
    assert it == g["vel_it"][0]                                                # identical iteration count
    assert relmax(x, g["vel_x"]) <= TOL_CONVERGED
    # pressure: constant null space (no outflow side)
    assert bool(g["ifvcor"][0])
    nek.h1mg_setup(fbc, case.xm1, case.ym1, case.zm1, case.vertex, E, True)
    z, r = np.zeros(n), g["rhs"].copy()
    nek.h1mg_solve(z, r, False)
    assert np.array_equal(r, g["rhs_out"]) and relmax(z, g["z"]) <= TOL_FIELD
    tol = float(g["tol"][0])
    nek.set_pressure_state(g["pmask"], g["binvm1"], tol, tol, True, E)
    res = g["b"].copy()
    it = nek.hmh_gmres(res, np.ones(n), np.zeros(n), g["vmult"], 100)
    assert it == g["it"][0] and relmax(res, g["x"]) <= TOL_FIELD
    res = g["b"].copy()
    it = nek.hmh_flex_cg(res, np.ones(n), np.zeros(n), g["vmult"], 100)
    assert it == g["it_fcg"][0] and relmax(res, g["x_fcg"]) <= 1e-9


def test_full_turbchannel_mesh_pressure_solve_against_the_reference(nek):
    """The complete mesh of examples/turbChannel (1536 elements, 786,432 points): h1mg_solve, then hmh_gmres (56 iterations:
    one GMRES(30) restart) and hmh_flex_cg (62) -- iteration counts identical to the reference's, fields at the 4096 sampled
    positions and their 2-norms within 1e-10."""
    g = dict(np.load(refcases.GOLDEN_CHANNEL_FULL))
    case = refcases.channel_case(refcases.CHANNEL_FULL_DIMS)
    E, n, idx = case.nel, case.n, g["idx"]
    geo = case.geom()                      # the oracle's factors are the reference's bit for bit (tests/test_ref_pins.py)
    _register(nek, case, geo[:6], geo[6], case.binv(), g["volvm1"][0], case.z, case.w, case.D)
    pmask = np.ones(n)
    rhs, b = refcases.pressure_inputs(case, pmask)
    for k, v in (("rhs", rhs), ("b", b)):
        assert np.array_equal(v[idx], g[k + "_s"])                             # the reference run's inputs
    nek.h1mg_setup(refcases.channel_fbc(case), case.xm1, case.ym1, case.zm1, case.vertex, E, True)
    z, r = np.zeros(n), rhs.copy()
    nek.h1mg_solve(z, r, False)
    assert np.array_equal(r[idx], g["rhs_out_s"])
    assert np.abs(z[idx] - g["z_s"]).max() <= TOL_FIELD * g["z_max"][0]
    assert abs(np.sqrt(np.sum(z * z)) - g["z_l2"][0]) <= TOL_FIELD * g["z_l2"][0]
    tol = float(g["tol"][0])
    nek.set_pressure_state(pmask, case.binv(), tol, tol, True, E)
    for solver, key, itkey, ftol in ((nek.hmh_gmres, "x", "it", TOL_FIELD), (nek.hmh_flex_cg, "x_fcg", "it_fcg", 1e-9)):
        res = b.copy()
        it = solver(res, np.ones(n), np.zeros(n), case.mult, 100)
        assert it == g[itkey][0], (key, it, g[itkey])
        assert np.abs(res[idx] - g[key + "_s"]).max() <= ftol * g[key + "_max"][0], key
        assert abs(np.sqrt(np.sum(res * res)) - g[key + "_l2"][0]) <= ftol * g[key + "_l2"][0], key


def test_cggo_history_against_the_reference_lanczos_tridiagonal(nek):
    """The reference's own record of a cggo solve -- the Lanczos tridiagonal in common /tdarray/ (hmholtz.f:808-815), i.e.
    every alpha and beta -- against the device's (rtz1, rho) history, over the window in which CG has not yet amplified
    last-bit differences past 1e-9 (tests/test_gpu_parity.py explains the window); the count (101) must be identical."""
    import ctypes as C
    from nek5000_b200 import lib
    from nek5000_b200._lib import check
    from nek5000_b200.nek import DevArray
    g, case = G["core"], refcases.case_of("core")
    n = case.n
    _register(nek, case, [g[f"g{i}m1"] for i in range(1, 7)], g["bm1"], g["binvm1"], g["volvm1"][0], g["zgm1"], g["wxm1"], g["dxm1"])
    d = [DevArray.from_host(a) for a in (np.zeros(n), g["cggo_f"], g["h1"], g["h2"], g["v1mask"], g["vmult"], g["binvm1"])]
    k = int(g["cggo_it"][0])
    hist, it = np.zeros(3 * (500 + 2)), C.c_int(0)              # 3 * (maxit + 2) doubles (include/nekb200.h)
    check(lib().nekb_cggo_dev(*[a.ptr for a in d], 1, 1e-6, 500, C.byref(it), hist.ctypes.data))
    assert it.value == k
    h = hist.reshape(-1, 3)
    rtz, rho = h[:, 0], h[:, 2]
    win = 40
    beta = np.zeros(win)
    beta[1:] = rtz[1:win] / rtz[:win - 1]
    diag = np.array([rho[0] / rtz[0]] + [(beta[i] ** 2 * rho[i - 1] + rho[i]) / rtz[i] for i in range(1, win)])
    upper = np.array([-beta[i] * rho[i - 1] / np.sqrt(rtz[i - 1] * rtz[i]) for i in range(1, win)])
    assert np.all(np.abs(diag - g["cggo_diagt"][:win]) <= 1e-9 * np.abs(g["cggo_diagt"][:win]))
    assert np.all(np.abs(upper - g["cggo_upper"][:win - 1]) <= 1e-9 * np.abs(g["cggo_upper"][:win - 1]))


@pytest.mark.parametrize("null_space", [False, True])
def test_crs_facade_against_a_dense_solve(nek, null_space):
    """core/fcrs.c crs_setup / crs_solve (XXT slot) on the vertex mesh of a 4 x 3 x 2 box: ids = vertex numbers (0 on a
    Dirichlet side unless the null space is kept), element matrices with the constant in their kernel, COO indices as
    set_mat_ij builds them.  Against numpy: x = Q A^-1 Q^T b on the distinct dofs, 0 on ignored ones; with the null space the
    mean-free solution of a consistent right-hand side (crs_xxt.c:951-960)."""
    import oracle
    case = oracle.Case(4, 3, 2, nx=2)
    E = case.nel
    ids = case.vertex.reshape(E, 8).copy()
    rng = np.random.default_rng(5)
    if not null_space:
        ids[np.isin(ids, np.unique(ids)[:12])] = 0                # twelve vertices play the Dirichlet side
    B = rng.standard_normal((E, 8, 8))
    Ae = np.einsum("eik,ejk->eij", B, B)
    P = np.eye(8) - np.full((8, 8), 1.0 / 8)
    Ae = np.einsum("ik,ekl,lj->eij", P, Ae, P)                    # A_e 1 = 0: the assembled operator keeps the constant null space
    loc = np.arange(8 * E).reshape(E, 8)
    Ai = np.repeat(loc[:, :, None], 8, axis=2)                    # ia(i,j,e) = (e-1) n + i - 1, ja(i,j,e) = (e-1) n + j - 1
    Aj = np.repeat(loc[:, None, :], 8, axis=1)
    h = nek.crs_setup(ids.ravel(), Ai.ravel(), Aj.ravel(), Ae.ravel(), null_space)
    gids = np.unique(ids[ids != 0])
    dof = np.searchsorted(gids, ids.ravel())
    keep = ids.ravel() != 0
    nc = len(gids)
    A = np.zeros((nc, nc))
    np.add.at(A, (dof[Ai.ravel()][keep[Ai.ravel()] & keep[Aj.ravel()]], dof[Aj.ravel()][keep[Ai.ravel()] & keep[Aj.ravel()]]),
              Ae.ravel()[keep[Ai.ravel()] & keep[Aj.ravel()]])
    b = rng.standard_normal(8 * E)
    g = np.zeros(nc)
    np.add.at(g, dof[keep], b[keep])
    if null_space:
        b[keep] -= g.sum() / keep.sum()                           # consistent: the assembled right-hand side sums to zero
        g = np.zeros(nc)
        np.add.at(g, dof[keep], b[keep])
        y = np.linalg.lstsq(A, g, rcond=None)[0]
        y -= y.mean()
    else:
        y = np.linalg.solve(A, g)
    want = np.where(keep, y[np.minimum(dof, nc - 1)], 0.0)
    got = nek.crs_solve(h, b)
    assert np.array_equal(got[~keep], np.zeros((~keep).sum()))
    assert relmax(got, want) <= 1e-9
    nek.crs_free(h)


@pytest.mark.parametrize("m,nmax,omega_p,iters", [(20, 4096, 0.0, 27), (20, 4096, 0.66, 20), (8, 4096, 0.0, 1), (32, 800, 0.0, None),
                                                  (32, 800, 0.66, None)])
@pytest.mark.parametrize("coop", ["1", "0"])
def test_crs_amg_device_cycle_against_numpy(nek, m, nmax, omega_p, iters, coop, monkeypatch):
    """coop = "1": the whole solve in one cooperative launch (amg_pcg_coop_kernel, default); "0": one launch per operation.
    CG + V(1,1) aggregation cycle on the device against the same algorithm in numpy on the same (library-built) levels:
    same iteration count (27 at 8820 vertices, 20 with the smoothed prolongation: the prototype's numbers), solution to 1e-10, residual at the requested 1e-13."""
    import ctypes as C
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "scripts"))
    import proto_coarse_amg as proto
    import scipy.sparse as sp
    from nek5000_b200 import lib
    from nek5000_b200._lib import check
    from nek5000_b200.nek import DevArray
    monkeypatch.setenv("NEKB_CRS_AMG_COOP", coop)
    A = proto.q1_stiffness(m)
    n = A.shape[0]
    co = A.tocoo()
    lv = nek.crs_amg_build_host(n, co.row, co.col, co.data, nmax=nmax, theta=0.02, omega_p=omega_p)
    check(lib().nekb_crs_amg_upload(0.7))
    mats = [sp.csr_matrix((l["val"], l["col"], l["rowptr"]), shape=(l["n"], l["n"])) for l in lv]
    Ps = [sp.csr_matrix((l["p_val"], l["p_col"], l["p_rowptr"]), shape=(l["n"], lv[k + 1]["n"])) for k, l in enumerate(lv[:-1])]
    Ainv = np.linalg.inv(mats[-1].toarray())

    def cycle(b, l=0):
        if l == len(mats) - 1:
            return Ainv @ b
        dj = 0.7 / mats[l].diagonal()
        x = dj * b
        x = x + Ps[l] @ cycle(Ps[l].T @ (b - mats[l] @ x), l + 1)
        return x + dj * (b - mats[l] @ x)

    xe = np.random.default_rng(0).standard_normal(n)
    b = A @ xe
    xref, itref = proto.pcg(A, b, cycle)
    bd, xd = DevArray.from_host(b), DevArray(n)
    it = C.c_int(0)
    check(lib().nekb_crs_amg_solve_dev(xd.ptr, bd.ptr, 1e-13, 500, C.byref(it)))
    x = xd.to_host()
    # 1e-13 is the rounding floor of the recurrence: FMA contraction and the reduction order may move the exit by an iteration
    assert abs(it.value - itref) <= 2 and (iters is None or itref == iters)
    assert relmax(x, xref) <= 1e-10 and relmax(x, xe) <= 1e-9
    assert np.linalg.norm(A @ x - b) <= 5e-13 * np.linalg.norm(b)


@pytest.mark.parametrize("name", ["ethier", "channel"])
def test_h1mg_solve_with_the_aggregation_coarse_solver(nek, name, monkeypatch):
    """h1mg_solve / hmh_gmres with the coarse problem solved by CG over the aggregation hierarchy instead of the dense inverse
    (forced on these small meshes: dense path off, coarsest level <= 8 unknowns): same golden fields and iteration counts."""
    monkeypatch.setenv("NEKB_CRS_DENSE", "0")
    monkeypatch.setenv("NEKB_CRS_AMG", "1")
    monkeypatch.setenv("NEKB_CRS_AMG_NMAX", "8")
    g = G[name]
    case = refcases.channel_case() if name == "channel" else refcases.ethier_case()
    fbc = refcases.channel_fbc(case) if name == "channel" else hsmg.box_fbc(case, (2,) * 6)
    E, n = case.nel, case.n
    _register(nek, case, [g[f"g{i}m1"] for i in range(1, 7)], g["bm1"], g["binvm1"], g["volvm1"][0], g["zgm1"], g["wxm1"], g["dxm1"])
    nek.h1mg_setup(fbc, case.xm1, case.ym1, case.zm1, case.vertex, E, True)
    z, r = np.zeros(n), g["rhs"].copy()
    nek.h1mg_solve(z, r, False)
    assert relmax(z, g["z"]) <= TOL_FIELD
    assert 2 <= nek.h1mg_info()["crs_iters"] <= 40
    tol = float(g["tol"][0])
    nek.set_pressure_state(g["pmask"], g["binvm1"], tol, tol, True, E)
    res = g["b"].copy()
    it = nek.hmh_gmres(res, np.ones(n), np.zeros(n), g["vmult"], 100)
    assert it == g["it"][0] and relmax(res, g["x"]) <= TOL_FIELD


@pytest.mark.parametrize("nx", [4, 6, 10])
def test_h1mg_and_gmres_at_other_orders_against_the_reference(nx):
    """lx1 = 4 (two multigrid levels), 6 and 10: h1mg_solve, hmh_gmres and hmh_flex_cg directly against the reference's output
    (golden h1mg_lx{nx}); geometry from the oracle, which is the reference's bit for bit."""
    from nek5000_b200 import nek
    g, case = G[f"h1mg_lx{nx}"], refcases.case_of("core", nx)
    E, n = case.nel, case.n
    nek.finalize()
    nek.init(0, nx, 3)
    try:
        geo = case.geom()
        nek.set_nel(E, E)
        nek.set_gll(case.z, case.w)
        nek.set_dxyz(case.D, np.ascontiguousarray(case.D.T))
        nek.set_geom(*geo[:6], geo[6])
        nek.set_ifdfrm(None)
        h, _ = nek.setupds(nx, E, case.vertex)
        nek.set_ifield(1)
        nek.set_field_handle(1, h)
        nek.set_step_info(1, float(g["volvm1"][0]))
        nek.set_binv(case.binv())
        nek.h1mg_setup(refcases.fbc_of("core", case), case.xm1, case.ym1, case.zm1, case.vertex, E, False)
        z, r = np.zeros(n), g["rhs"].copy()
        nek.h1mg_solve(z, r, False)
        assert np.array_equal(r, g["rhs_out"]) and relmax(z, g["z"]) <= TOL_FIELD
        tol = float(g["tol"][0])
        nek.set_pressure_state(g["pmask"], case.binv(), tol, tol, False, E)
        res = g["b"].copy()
        assert nek.hmh_gmres(res, np.ones(n), np.zeros(n), case.mult, 100) == g["it"][0] and relmax(res, g["x"]) <= TOL_FIELD
        res = g["b"].copy()
        assert nek.hmh_flex_cg(res, np.ones(n), np.zeros(n), case.mult, 100) == g["it_fcg"][0] and relmax(res, g["x_fcg"]) <= 1e-9
    finally:
        nek.finalize()


def test_hsolve_pres_on_the_channel_mesh_as_turbchannel_par_runs_it(nek):
    """BASELINE config 5 with turbChannel.par's settings (residualTol 1e-4, residualProj = yes): hsolve('PRES') -> project1 ->
    hmhzpf -> cggo('PRES') -> hmh_gmres with h1mg_solve -> project2 on the channel mesh with the constant null space; the
    reference needs 10, 8, 7, 2 iterations for the four successive solves."""
    g, case = G["hsolve_pres_channel"], refcases.channel_case()
    gc = G["channel"]
    E, n = case.nel, case.n
    _register(nek, case, [gc[f"g{i}m1"] for i in range(1, 7)], gc["bm1"], gc["binvm1"], gc["volvm1"][0], gc["zgm1"], gc["wxm1"], gc["dxm1"])
    nek.h1mg_setup(refcases.channel_fbc(case), case.xm1, case.ym1, case.zm1, case.vertex, E, True)
    nek.set_pressure_state(g["mask"], g["binvm1"], 1e-4, 1e-4, True, E)
    nek.set_param(22, 0.0), nek.set_param(42, 0.0), nek.set_param(93, 20.0), nek.set_param(95, 5.0)
    nek.projection_reset()
    napprox = np.zeros(10, dtype=np.int32)
    for k, (rhs, h1, h2, istep) in enumerate(refcases.hsolve_inputs(case, pres=True, consistent=True)):
        nek.set_step_info(istep, float(g["volvm1"][0]))
        u, r = np.zeros(n), rhs.copy()
        it = nek.hsolve("PRES", u, r, h1, h2, g["mask"], g["vmult"], 1, 1e-4, 200, 1, None, napprox, g["binvm1"])
        assert napprox[1] == g["m"][k]
        refcases.count_or_margin(it, int(g["its"][k]), nek.last_history()[:, 0], g[f"res{k}"], float(g[f"restol{k}"][0]),
                                 cap=1e-7, what=f"channel hsolve('PRES') call {k}", gmres=True)
        assert relmax(u, g[f"u{k}"]) <= 1e-5, k


def test_ethier_par_velocity_solve_where_chktcg1_bites(nek):
    """ethier.par's literal velocity solve (viscosity 0.1, dt 1e-4 / bdf3, residualTol 1e-12): hmholtz's chktcg1 raises the
    tolerance to 1.8e-9; the reference stops after 7 iterations."""
    g, case = G["ethier"], refcases.ethier_case()
    n = case.n
    _register(nek, case, [g[f"g{i}m1"] for i in range(1, 7)], g["bm1"], g["binvm1"], g["volvm1"][0], g["zgm1"], g["wxm1"], g["dxm1"])
    nek.set_step_info(10, float(g["volvm1"][0]))
    nek.set_param(22, 0.0)
    x, rhs = np.zeros(n), g["par_rhs"].copy()
    it = nek.hmholtz("VELX", x, rhs, np.full(n, g["par_h1"][0]), np.full(n, g["par_h2"][0]), g["v1mask"], g["vmult"], 1, 1e-12, 200, 1)
    assert it == g["par_it"][0] == 7 and relmax(x, g["par_x"]) <= TOL_CONVERGED
