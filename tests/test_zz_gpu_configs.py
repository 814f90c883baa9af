"""GPU parity on the meshes BASELINE.json's configs name, against the REFERENCE ITSELF (oracle/_ref, see tests/refcases.py):

  * config 2 -- short_tests/ethier: the 27-element box on [-1,1]^3 with 'v  ' on every side;
  * config 5 -- examples/turbChannel: periodic in x and z, walls in y, vertices stretched by the case's usrdat; a 4^3-element
    cut with full golden fields, and the complete 16 x 12 x 8 = 1536-element mesh with sampled golden fields.

Tolerances are the north star's: identical CG / GMRES iteration counts, fields <= 1e-10 relative (a solve run to
convergence is compared at the accuracy it was asked for, as in tests/test_gpu_golden.py).  The file sorts after the other
GPU suites on purpose: these are the largest cases.
"""
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
