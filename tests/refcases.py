"""Golden cases run through the REFERENCE ITSELF (oracle/_ref: /root/reference's own Fortran statements, transpiled by
oracle/f77c.py and compiled with gcc -- see oracle/ref_build.py).  TEST INFRASTRUCTURE.

`reference(name)` runs one case in the transpiled reference and returns {key: array} holding the inputs it drew and the
outputs the reference produced; `tests/golden/gen_ref_golden.py` stores them in `tests/golden/ref_golden.npz`, which is
what travels: the CPU suite holds the oracle to it and the GPU suite holds the CUDA path to it.

Each case names the reference routines it executes (file:line in the docstring of its function).
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

import oracle
from oracle import hsmg

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_golden.npz")

# name -> (dims, dirichlet sides of the velocity (x-,x+,y-,y+,z-,z+; 0 = outflow 'O  '), deformation)
MESH = {
    "core": ((3, 2, 2), (1, 1, 1, 1, 1, 0), 0.05),
    "neumann": ((3, 2, 2), (1, 1, 1, 1, 1, 1), 0.05),
    "fdm": ((3, 3, 2), (1, 1, 0, 0, 1, 1), 0.02),
    "pnpn2": ((3, 2, 2), (1, 1, 1, 1, 1, 0), 0.0),
    "ophinv": ((3, 2, 2), (1, 2, 1, 2, 1, 0), 0.05),
    "eop": ((3, 2, 2), (1, 2, 1, 1, 1, 0), 0.05),          # Pn-Pn-2 E operator: deformed, wall / symmetry / outflow sides      # 2 = 'SYM': the three velocity masks differ
}


def case_of(name, nx=8):
    dims, dirich, deform = MESH[name]
    return oracle.Case(*dims, nx=nx, dirichlet=dirich, deform=deform)


def fbc_of(name, case, bsym=2):
    """get_fast_bc codes (core/fast3d.f:802-877) on the box sides: 'v  ' -> 2 (Neumann for p), 'O  ' -> 1 (Dirichlet),
    'SYM' -> bsym (2 for the multigrid levels, hsmg.f:725; 3 for gen_fast, fast3d.f:38)."""
    return hsmg.box_fbc(case, tuple({0: 1, 1: 2, 2: bsym}[int(d)] for d in MESH[name][1]))


def _ref(case, **kw):
    from oracle.ref import RefCase
    return RefCase(case, **kw)


class _Logged:
    """Runs reference calls with the reference's own per-iteration log lines switched on (nio = 0, ifprint, param(74); the
    values are captured by oracle/ref_stubs.c f77_trace instead of being printed) and returns what was logged:
    cggo (core/hmholtz.f:770-773): istep, iter, rbn2, h1(1), tol, h2(1), ifmcor per executed check;
    hmh_gmres (core/gmres.f:496-498): iter, tolpss, rnorm, div0, ratio, istep per iteration."""

    def __init__(self, R):
        self.R = R

    def __enter__(self):
        R = self.R
        self.saved = (int(R.get("nio")), int(R.get("ifprint")), float(R.var("param")[73]))
        R.set("nio", 0), R.set("ifprint", 1)
        R.var("param")[73] = 1.0
        R.trace(True)
        return self

    def __exit__(self, *exc):
        R = self.R
        self.records = R.trace_records()
        R.trace(False)
        R.set("nio", self.saved[0]), R.set("ifprint", self.saved[1])
        R.var("param")[73] = self.saved[2]

    def cggo(self):
        """(rbn2 per check, tol) of the cggo calls logged, one pair per call (a call starts at iter = 1)."""
        calls = []
        for u, v in self.records:
            if u == "cggo" and len(v) == 7:
                if int(v[1]) == 1:
                    calls.append(([], v[4]))
                calls[-1][0].append(v[2])
                calls[-1] = (calls[-1][0], v[4])
        return [(np.array(h), float(t)) for h, t in calls]

    def gmres(self):
        calls = []
        for u, v in self.records:
            if u == "hmh_gmres" and len(v) == 6:
                if int(v[0]) == 1:
                    calls.append(([], v[1]))
                calls[-1][0].append(v[2])
        return [(np.array(h), float(t)) for h, t in calls]


def ulp_perturbed(a, seed=99):
    """`a` with every entry moved by one unit of rounding (relative 2^-52, random sign): the smallest change of the input a
    floating-point computation can see.  The reference run on such an input shows how far the reference's OWN iterates move
    under rounding-level changes -- the yardstick for an iteration count that differs by one (tests: count_or_margin)."""
    sgn = np.where(np.random.default_rng(seed).random(a.shape) < 0.5, -1.0, 1.0)
    return a * (1.0 + 2.0 ** -52 * sgn)


def count_or_margin(it, it_ref, res, res_ref, tol, res_ref_pert=None, slack=1000.0, cap=1e-6, what="", gmres=False):
    """The north star asks for IDENTICAL iteration counts.  This passes when they are.  When they differ, it passes only for a
    difference of one that is a proven rounding-margin event, and it says so in the assertion text otherwise:

      * k = min(it, it_ref) is the exit check that went the other way (cggo: check k+1 leaves with niterhm = k,
        core/hmholtz.f:778-779; gmres = True: the test after iteration k, core/gmres.f:502, i.e. entry k-1 of the histories);
        `res` / `res_ref` are the residual norms that check compares with `tol` (what the reference logs per iteration,
        core/hmholtz.f:770-773, core/gmres.f:496-498);
      * with `res_ref_pert` (the reference run again on an input moved by ONE unit of rounding): the distance of the
        reference's residual from tol at check k must be within `slack` x what the reference's own history moved under that
        perturbation (envelope up to k) -- i.e. the reference itself flips under rounding-level changes of this size -- and
        the device history must stay within `slack` x that envelope of the reference's at every earlier check;
      * without it: the device history must agree with the reference's to `cap` (relative) at every check up to k, which
        bounds the distance of the reference's residual from tol by the same number (the two residuals straddle tol)."""
    if it == it_ref:
        return
    assert abs(it - it_ref) == 1, f"{what}: iteration count {it} vs the reference's {it_ref}"
    k = min(it, it_ref) - (1 if gmres else 0)
    assert len(res) > k and len(res_ref) > k, f"{what}: histories too short ({len(res)}, {len(res_ref)}) for check {k}"
    rel = np.abs(np.asarray(res[:k + 1]) - res_ref[:k + 1]) / res_ref[:k + 1]
    margin = abs(res_ref[k] - tol) / tol
    assert (res[k] <= tol) != (res_ref[k] <= tol), f"{what}: counts {it}/{it_ref} differ but check {k} agrees: {res[k]}, {res_ref[k]}, tol {tol}"
    if res_ref_pert is not None:
        m = min(len(res_ref_pert), k + 1)
        env = np.maximum.accumulate(np.abs(res_ref[:m] - res_ref_pert[:m]) / res_ref[:m])
        env = np.concatenate([env, np.full(k + 1 - m, env[-1])])
        assert margin <= slack * env[k], (f"{what}: count {it} vs {it_ref}: the reference's residual is {margin:.2e} (relative) away from "
                                          f"tol at check {k}, its own rounding sensitivity there is {env[k]:.2e}")
        bad = rel > slack * np.maximum(env, 1e-15)
        assert not bad.any(), f"{what}: history departs from the reference's beyond its rounding sensitivity at checks {np.flatnonzero(bad)[:5]}: {rel[bad][:5]} vs {env[bad][:5]}"
    else:
        assert rel.max() <= cap, f"{what}: count {it} vs {it_ref} and the residual histories differ by {rel.max():.2e} > {cap:.0e}"
        assert margin <= rel[k] * (1 + 1e-12)


# --------------------------------------------------------------------------------------------------- reference runs
def ref_core(nx=8):
    """setupds/setvert3d (navier8.f:2004-2360), geom1/geom2/setinvm (coef.f:555-784), bcmask (bdry.f:317+), axhelm
    (hmholtz.f:72-259), setprec (:380-524), dssum/dsop (dssum.f:33-161), cggo (:611-846), hmholtz (:2-69), and the BP5
    driver bp5/cggos/geodatstd/rand_fld_h1 (examples/bp5/bp5.usr:324-395,797-899,623-699; navier5.f:2650-2700)."""
    case = case_of("core", nx)
    rc = _ref(case)
    R, n = rc.R, case.n
    out = dict(glo_num=R.var("glo_num").ravel(order="F")[:n].copy(), vmult=rc.fld("vmult"), bm1=rc.fld("bm1"), binvm1=rc.fld("binvm1"),
               v1mask=rc.fld("v1mask"), pmask=rc.fld("pmask"), volvm1=np.array([R.get("volvm1")]),
               zgm1=R.var("zgm1")[:, 0].copy(), wxm1=R.var("wxm1").copy(), dxm1=R.var("dxm1").copy(order="C"))
    for i in range(1, 7):
        out[f"g{i}m1"] = rc.fld(f"g{i}m1")
    rng = np.random.default_rng(1)
    u, h1, h2 = rng.standard_normal(n), 1.0 + rng.random(n), rng.random(n)
    out.update(u=u, h1=h1, h2=h2)
    au = np.zeros(n)
    R.call("setfast", h1, h2, 1)
    R.call("axhelm", au, u, h1, h2, 1, 1)
    out["axhelm"] = au
    au0, one, zero = np.zeros(n), np.ones(n), np.zeros(n)
    R.call("setfast", one, zero, 1)                               # Poisson: h1 = 1, h2 = 0 (ifh2 is set by setfast)
    R.call("axhelm", au0, u, one, zero, 1, 1)
    out["axhelm_poisson"] = au0
    dp = np.zeros(n)
    R.call("setfast", h1, h2, 1)
    R.call("setprec", dp, h1, h2, 1, 1)
    out["setprec"] = dp
    for op, key in (("+  ", "dsop_add"), ("*  ", "dsop_mul"), ("m  ", "dsop_min"), ("M  ", "dsop_max")):
        v = u.copy()
        R.call("dsop", v, op, nx, nx, nx)
        out[key] = v
    # cggo, Jacobi branch
    f = case.dssum(rng.standard_normal(n) * case.bm1()) * case.mask
    out["cggo_f"] = f
    for key, tin, maxit in (("cggo20", 1e-30, 20), ("cggo", 1e-6, 500)):
        x = np.zeros(n)
        R.set("kfldfdm", -1), R.set("ifsolv", 0), R.set("istep", 1)
        R.call("cggo", x, f.copy(), h1, h2, case.mask, case.mult, 1, tin, maxit, 1, rc.fld("binvm1"), "VELX")
        out[key + "_x"], out[key + "_it"] = x, np.array([R.get("niterhm")])
    # the Lanczos tridiagonal cggo leaves in common /tdarray/ (hmholtz.f:808-815): diag_k = (beta_k^2 rho_{k-1} + rho_k) /
    # rtz1_k, upper_{k-1} = -beta_k rho_{k-1} / sqrt(rtz2 rtz1) -- the whole alpha / beta history of the converged solve
    nit = int(out["cggo_it"][0])
    out["cggo_diagt"], out["cggo_upper"] = R.var("diagt")[:nit].copy(), R.var("upper")[:nit - 1].copy()
    # hmholtz wrapper on an un-assembled right-hand side
    rhs = case.bm1() * rng.standard_normal(n)
    out["hmh_rhs"] = rhs
    x, r = np.zeros(n), rhs.copy()
    R.var("param")[21] = 0.0
    R.set("ifsolv", 0)
    R.call("hmholtz", "VELX", x, r, h1, h2, case.mask, case.mult, 1, 1e-7, 300, 1)
    out["hmh_x"], out["hmh_it"], out["hmh_rhs_out"] = x, np.array([R.get("niterhm")]), r
    # BP5 (40 fixed iterations)
    R.var("uparam")[0:3] = (-1e-8, 40, 1)
    R.call("bp5")
    v = lambda nm: R.var(nm, "bp5").ravel(order="F")
    out.update(bp5_gf=v("gf")[:6 * n].copy(), bp5_e1=v("e1")[:n].copy(), bp5_r1=v("r1")[:n].copy(), bp5_u1=v("u1")[:n].copy())
    return out


def ref_core_lx6():
    """The same routines at lx1 = 6 (the library's generic, non-TMA kernels; a second polynomial order for the oracle)."""
    keep = ("glo_num", "vmult", "bm1", "binvm1", "v1mask", "volvm1", "zgm1", "wxm1", "dxm1", "g1m1", "g2m1", "g3m1", "g4m1", "g5m1",
            "g6m1", "u", "h1", "h2", "axhelm", "axhelm_poisson", "setprec", "dsop_add", "cggo_f", "cggo20_x", "cggo20_it", "cggo_x",
            "cggo_it", "cggo_diagt", "cggo_upper", "bp5_gf", "bp5_e1", "bp5_r1", "bp5_u1")
    full = ref_core(6)
    return {k: full[k] for k in keep}


def ref_bp5():
    """The BP5 driver on the all-Dirichlet box (every side 'v  ', the benchmark's own boundary conditions,
    examples/bp5/genbox.in): bp5.usr:324-395 with 40 fixed iterations."""
    case = case_of("neumann")
    rc = _ref(case)
    R, n = rc.R, case.n
    R.var("uparam")[0:3] = (-1e-8, 40, 1)
    R.call("bp5")
    v = lambda nm: R.var(nm, "bp5").ravel(order="F")
    return dict(glo_num=R.var("glo_num").ravel(order="F")[:n].copy(), gf=v("gf")[:6 * n].copy(), e1=v("e1")[:n].copy(),
                r1=v("r1")[:n].copy(), u1=v("u1")[:n].copy())


def _pressure(name, nx=8, pcg=False):
    return _pressure_case(case_of(name, nx), pcg=pcg)


def channel_case(dims=(4, 4, 4), nx=8):
    """The mesh of examples/turbChannel (BASELINE config 5) at a size the 64-element reference build holds: genbox box that
    is periodic in x and z with walls in y (turbChannel.box), element vertices moved by the case's usrdat
    (turbChannel.usr:334-358: x scaled to XLEN = 2 pi, z to ZLEN = pi, y = tanh(BETAM (2y-1)) / tanh(BETAM), BETAM = 2.4)
    before the GLL points are generated."""
    def usrdat(xc, yc, zc):
        xs, ys, zs = 2 * np.pi / (xc.max() - xc.min()), 1.0 / (yc.max() - yc.min()), np.pi / (zc.max() - zc.min())
        return xs * xc, np.tanh(2.4 * (2 * (ys * yc) - 1)) / np.tanh(2.4), zs * zc
    return oracle.Case(*dims, nx=nx, periodic=(1, 0, 1), dirichlet=(1, 1, 1, 1, 1, 1), rescale=False, vertex_map=usrdat)


def channel_fbc(case):
    """get_fast_bc codes of the channel: periodic sides 0, walls 2 (Neumann for the pressure)."""
    return hsmg.box_fbc(case, (0, 0, 2, 2, 0, 0))


def ethier_case(nx=8):
    """The box of short_tests/ethier (BASELINE config 2; ethier.box: 3 x 3 x 3 elements, every side 'v  '), rescaled to
    [-1, 1]^3 by the case's usrdat2 (ethier.usr:220-228)."""
    return oracle.Case(3, 3, 3, nx=nx, lo=(-1.0, -1.0, -1.0), hi=(1.0, 1.0, 1.0), rescale=(-1.0, 1.0))


def geometry_of(rc):
    """What the device side registers for a case, as the reference computed it."""
    R = rc.R
    out = dict(vmult=rc.fld("vmult"), bm1=rc.fld("bm1"), binvm1=rc.fld("binvm1"), v1mask=rc.fld("v1mask"),
               zgm1=R.var("zgm1")[:, 0].copy(), wxm1=R.var("wxm1").copy(), dxm1=R.var("dxm1").copy(order="C"))
    for i in range(1, 7):
        out[f"g{i}m1"] = rc.fld(f"g{i}m1")
    return out


def pressure_inputs(case, pmask):
    """Right-hand sides in the range of the operator (A times a continuous field): with the all-Neumann null space an
    inconsistent coarse rhs has no meaningful answer -- XXT pins a different unknown for every rank count
    (crs_xxt.c:893-899, 926-949) -- and the residuals the preconditioner sees in hmh_gmres are consistent.  Returns the
    rhs of the h1mg_solve call and of the hmh_gmres / hmh_flex_cg calls."""
    n = case.n
    rng = np.random.default_rng(3)
    h1, h2 = np.ones(n), np.zeros(n)
    xr = case.dssum(rng.standard_normal(n)) * case.mult * pmask
    rhs = case.dssum(case.axhelm(xr, h1, h2)) * pmask
    xe = case.dssum(rng.standard_normal(n)) * case.mult * pmask
    b = case.dssum(case.axhelm(xe, h1, h2)) * pmask
    return rhs, b


def _pressure_case(case, with_geometry=False, capped=0, pcg=False):
    """set_overlap -> hsmg_setup/h1mg_setup/set_up_h1_crs (navier6.f:29-101, hsmg.f:22-47,2234-2270, navier8.f:83-233),
    h1mg_solve (hsmg.f:1855-1949) and hmh_gmres (gmres.f:304-545) incl. chktcg1 and ortho."""
    rc = _ref(case)
    R, n = rc.R, case.n
    R.set("ifmgrid", 1)
    R.var("param")[[39, 40, 41, 42, 43]] = 0.0
    R.call("set_overlap")
    pmask = rc.fld("pmask")
    h1, h2 = np.ones(n), np.zeros(n)
    rhs, b = pressure_inputs(case, pmask)
    z, r = np.zeros(n), rhs.copy()
    R.call("h1mg_solve", z, r, False)
    tol = 1e-8
    R.var("param")[20] = tol
    R.set("tolps", tol), R.set("istep", 1)
    x, it = b.copy(), C.c_int(100)
    R.call("hmh_gmres", x, h1, h2, case.mult, it)
    # residual history of the last GMRES cycle as the reference's own Givens data holds it (gmres.f:486-493): after step k
    # rnorm_k = |s_k| rnorm_{k-1}, and the final rnorm = |gamma(j+1)| norm_fac
    j = it.value - 30 * ((it.value - 1) // 30)
    gm_s = np.abs(R.var("s_gmres")[:j]).copy()
    gm_last = np.array([abs(R.var("gamma_gmres")[j]) / np.sqrt(R.get("volvm1"))])
    xf, itf = b.copy(), C.c_int(100)
    R.call("hmh_flex_cg", xf, h1, h2, case.mult, itf)            # core/hmholtz.f:2164 (param(42) = 2)
    pcg_out = {}
    if pcg:
        # the plain PCG pressure solve (param(42) = 1): cggo('PRES') = Schwarz (fdm_h1, field ldim+1) + crs_solve_h1 + ortho
        # (core/hmholtz.f:660-846 with :710-712, :741-748); hmholtz sets kfldfdm and calls set_fdm_prec_h1A on first use (:22-50)
        R.var("param")[41] = 1.0
        R.call("set_fdm_prec_h1a")
        kt4 = R.var("ktype")[:, :, 4].copy(order="F")
        for key, rhs_p in (("pcg", b), ("pcg_pert", ulp_perturbed(b))):
            xp = np.zeros(n)
            R.set("ifsolv", 0), R.set("kfldfdm", 4), R.set("ifield", 1), R.set("istep", 1)
            with _Logged(R) as lg:
                R.call("cggo", xp, rhs_p.copy(), h1, h2, pmask, case.mult, 1, tol, 200, 1, rc.fld("binvm1"), "PRES")
            (hist_p, tol_p), = lg.cggo()
            pcg_out.update({f"x_{key}": xp, f"it_{key}": np.array([R.get("niterhm")]), f"{key}_rbn2": hist_p, "pcg_tol": np.array([tol_p])})
        pcg_out["ktype_pres"] = kt4[:case.nel].astype(np.int32)
        R.set("kfldfdm", -1)
        R.var("param")[41] = 0.0
    out = dict(pmask=pmask, rhs=rhs, rhs_out=r, z=z, b=b, x=x, it=np.array([it.value]), tol=np.array([tol]), **pcg_out,
               x_fcg=xf, it_fcg=np.array([itf.value]), gmres_s=gm_s, gmres_rnorm_last=gm_last,
               ifvcor=np.array([int(R.get("ifvcor"))]), volvm1=np.array([R.get("volvm1")]))
    if capped:                                                     # the same GMRES stopped after `capped` iterations
        xc, itc = b.copy(), C.c_int(capped)
        R.call("hmh_gmres", xc, h1, h2, case.mult, itc)
        out.update(x_capped=xc, it_capped=np.array([itc.value]))
    if with_geometry:
        out.update(geometry_of(rc))
        out.update(_velocity_solve(rc, case))
    return out


def _velocity_solve(rc, case):
    """One velocity Helmholtz solve of a time step on the same mesh: hmholtz('VELX') (hmholtz.f:2-69 -> cggo :611-846) with
    constant h1 = 0.01 (viscosity) and h2 = 500 (bd/dt) -- a mass-dominated balance of the kind a time step produces, not the
    literal values of ethier.par (viscosity 0.1, dt 1e-4, bdf3) -- on an un-assembled right-hand side, tolerance 1e-9."""
    R, n = rc.R, case.n
    rng = np.random.default_rng(11)
    h1, h2 = np.full(n, 0.01), np.full(n, 1.0 / 2e-3)
    rhs = rc.fld("bm1") * rng.standard_normal(n)
    x, r = np.zeros(n), rhs.copy()
    R.var("param")[21] = 0.0
    R.set("ifsolv", 0), R.set("kfldfdm", -1), R.set("istep", 1), R.set("ifield", 1)
    R.call("hmholtz", "VELX", x, r, h1, h2, rc.fld("v1mask"), rc.fld("vmult"), 1, 1e-9, 200, 1)
    return dict(vel_h1=h1[:1].copy(), vel_h2=h2[:1].copy(), vel_rhs=rhs, vel_x=x, vel_it=np.array([R.get("niterhm")]))


def ref_channel():
    """BASELINE config 5 (turbChannel mesh, 4 x 4 x 4 elements): pressure multigrid / GMRES / flexible CG with the constant
    null space on a periodic, wall-stretched box, and one velocity Helmholtz solve."""
    return _pressure_case(channel_case(), with_geometry=True, pcg=True)


GOLDEN_CHANNEL_FULL = os.path.join(os.path.dirname(GOLDEN), "ref_channel_full.npz")
CHANNEL_FULL_DIMS = (16, 12, 8)          # examples/turbChannel/turbChannel.box
CHANNEL_FULL_CAP = 8


def channel_full_samples(n):
    """4096 fixed sample positions of a field on the full mesh (the fields themselves are 6.3 MB each)."""
    return np.sort(np.random.default_rng(123).choice(n, 4096, replace=False))


def ref_channel_full():
    """BASELINE config 5 at the size of the reference's own files: the 16 x 12 x 8 = 1536-element mesh of
    examples/turbChannel/turbChannel.{box,re2} (tests/test_readers.py checks the generated vertices and vertex ids against
    those files), lx1 = 8, 786,432 grid points.  Needs the lelt = 1536 build of oracle/_ref (`python oracle/ref_build.py
    --lelt 1536`).  Stored: iteration counts, norms, and the fields at 4096 fixed positions; the inputs are regenerated on
    the test side (`pressure_inputs`) and identified by their norms and samples."""
    case = channel_case(CHANNEL_FULL_DIMS)
    g = _pressure_case(case, capped=CHANNEL_FULL_CAP)
    idx = channel_full_samples(case.n)
    out = dict(idx=idx.astype(np.int64), nel=np.array([case.nel]))
    for k, v in g.items():
        if v.size == case.n:
            out[k + "_s"], out[k + "_l2"], out[k + "_max"] = v[idx].copy(), np.array([np.sqrt(np.sum(v * v))]), np.array([np.abs(v).max()])
        else:
            out[k] = v
    return out


def ref_ethier():
    """BASELINE config 2 (ethier box, 27 elements on [-1,1]^3, all sides 'v  '): the same set, plus the velocity solve with the
    literal constants of short_tests/ethier/ethier.par: viscosity = -10 (h1 = 0.1), dt = 1e-4 with bdf3 (h2 = (11/6)/dt),
    [VELOCITY] residualTol = 1e-12."""
    case = ethier_case()
    out = _pressure_case(case, with_geometry=True)
    rc = _ref(case)
    R, n = rc.R, case.n
    rng = np.random.default_rng(12)
    h1, h2 = np.full(n, 0.1), np.full(n, (11.0 / 6.0) / 1e-4)
    rhs = rc.fld("bm1") * rng.standard_normal(n) * h2[0]              # the size of bd/dt * B u
    x, r = np.zeros(n), rhs.copy()
    R.var("param")[21] = 0.0
    R.set("ifsolv", 0), R.set("kfldfdm", -1), R.set("istep", 10), R.set("ifield", 1)
    R.call("hmholtz", "VELX", x, r, h1, h2, rc.fld("v1mask"), rc.fld("vmult"), 1, 1e-12, 200, 1)
    out.update(par_h1=h1[:1].copy(), par_h2=h2[:1].copy(), par_rhs=rhs, par_x=x, par_it=np.array([R.get("niterhm")]))
    return out


def ref_h1mg():
    return _pressure("core", pcg=True)


def ref_h1mg_neumann():
    return _pressure("neumann", pcg=True)


def ref_h1mg_lx6():
    """The same at lx1 = 6 (multigrid orders 1, 3, 5; core/hsmg.f:2272-2337)."""
    return _pressure("core", 6, pcg=True)


def ref_h1mg_lx4():
    """lx1 = 4: the two-level form (mg_h1_lmax = 2, orders 1, 3; core/hsmg.f:2293)."""
    return _pressure("core", 4, pcg=True)


def ref_h1mg_lx10():
    """lx1 = 10: orders 1, 3, 9."""
    return _pressure("core", 10, pcg=True)


def ref_periodic():
    """setupds / setvert3d (core/navier8.f:2004-2360) and the multiplicity on a box that is periodic in x and z: the vertex
    ids identify opposite sides, so faces, edges and corners wrap around."""
    case = oracle.Case(4, 3, 2, nx=8, periodic=(1, 0, 1))
    rc = _ref(case)
    R, n = rc.R, case.n
    u = np.random.default_rng(9).standard_normal(n)
    v = u.copy()
    R.call("dssum", v, 8, 8, 8)
    return dict(glo_num=R.var("glo_num").ravel(order="F")[:n].copy(), vmult=rc.fld("vmult"), v1mask=rc.fld("v1mask"), u=u, dssum=v)


def ref_fdm():
    """set_fdm_prec_h1A (hmholtz.f:1028-1272), set_fdm_prec_h1b (:1274-1358), fdm_h1 (:937-1026) and the Schwarz branch of
    cggo (:731-746)."""
    case = case_of("fdm")
    rc = _ref(case)
    R, n, E = rc.R, case.n, case.nel
    R.call("set_fdm_prec_h1a")
    kt = R.var("ktype")[:, :, 1].copy(order="F")
    rng = np.random.default_rng(4)
    h1, h2 = 1.0 + 0.3 * rng.random(n), 0.5 + 0.2 * rng.random(n)
    d = np.zeros(n)
    R.set("kfldfdm", 1)
    R.call("set_fdm_prec_h1b", d, h1, h2, E)
    r, z, w = rng.standard_normal(n), np.zeros(n), np.zeros(n)
    R.call("fdm_h1", z, r.copy(), d, case.mask, case.mult, E, kt, w)
    f = case.dssum(rng.standard_normal(n) * case.bm1()) * case.mask
    out = dict(ktype=kt[:E].astype(np.int32), dd=R.var("dd").T.copy(), elsize=R.var("elsize")[:, :E].copy(), h1=h1, h2=h2,
               d=d, r=r, z=z, f=f)
    for key, tin, maxit in (("cg20", 1e-30, 20), ("cg", 1e-8, 300)):
        x = np.zeros(n)
        R.set("ifsolv", 0), R.set("istep", 1), R.set("kfldfdm", 1)
        R.call("cggo", x, f.copy(), h1, h2, case.mask, case.mult, 1, tin, maxit, 1, rc.fld("binvm1"), "VELX")
        out[key + "_x"], out[key + "_it"] = x, np.array([R.get("niterhm")])
    # the converged solve once more with the reference's log captured, and on a right-hand side perturbed by one unit of
    # rounding: the Schwarz-preconditioned CG amplifies such changes, and the second history says by how much
    for key, rhs in (("cg_rbn2", f), ("cg_rbn2_pert", ulp_perturbed(f))):
        with _Logged(R) as lg:
            R.set("ifsolv", 0), R.set("istep", 1), R.set("kfldfdm", 1)
            R.call("cggo", np.zeros(n), rhs.copy(), h1, h2, case.mask, case.mult, 1, 1e-8, 300, 1, rc.fld("binvm1"), "VELX")
        (hist, tol), = lg.cggo()
        out[key], out[key.replace("rbn2", "it")], out["cg_tol"] = hist, np.array([R.get("niterhm")]), np.array([tol])
    R.set("kfldfdm", -1)
    return out


def ref_pnpn2():
    """Pn-Pn-2 (lx2 = lx1-2): set_overlap -> swap_lengths, gen_fast_spacing, gen_fast (fast3d.f), init_weight_op
    (fasts.f:310-413), hsmg_setup; hsmg_solve (hsmg.f:1376-1602) with local_solves_fdm (fasts.f:2-94)."""
    case = case_of("pnpn2")
    rc = _ref(case, lx2=6, ifsplit=False)
    R, E = rc.R, case.nel
    R.set("ifmgrid", 1)
    R.var("param")[[39, 40, 41, 42, 43]] = 0.0
    R.call("set_overlap")
    n2 = 6 ** 3 * E
    rng = np.random.default_rng(5)
    r, e = rng.standard_normal(n2), np.zeros(n2)
    R.call("hsmg_solve", e, r.copy())
    out = dict(r=r, e=e, df=R.var("df")[:, :E].T.copy())
    for nm in ("sr", "ss", "st"):
        out[nm] = R.var(nm)[:, :E].T.copy()
    return out


def ref_ophinv():
    """core/induct.f:1022-1090 ophinv (standard branch: three hsolve -> hmholtz -> cggo calls, core/navier4.f:562-634,
    core/hmholtz.f:2-69,611-846) on a box with wall, symmetry and outflow sides, variable h1 and h2; once to convergence and
    once with 15 fixed iterations.  The per-component iteration counts come from three separate hmholtz calls, which the
    reference's ophinv must (and does) reproduce exactly."""
    case = case_of("ophinv")
    rc = _ref(case)
    R, n = rc.R, case.n
    masks = [rc.fld(m) for m in ("v1mask", "v2mask", "v3mask")]
    rng = np.random.default_rng(7)
    # h2/h1 ~ 100: the conditioning of a velocity solve with a small time step (72 / 76 / 79 iterations to 1e-8).  Runs of
    # 130+ iterations exist too, but there the exit test (hmholtz.f:778) of any re-ordered summation sits within rounding
    # of the reference's and the count may move by one.
    h1, h2 = 1.0 + 0.3 * rng.random(n), 20.0 * (5.0 + rng.random(n))
    rhs = [case.bm1() * rng.standard_normal(n) for _ in range(3)]
    out = dict(v1mask=masks[0], v2mask=masks[1], v3mask=masks[2], vmult=rc.fld("vmult"), binvm1=rc.fld("binvm1"),
               volvm1=np.array([R.get("volvm1")]), h1=h1, h2=h2, i1=rhs[0], i2=rhs[1], i3=rhs[2])
    R.var("param")[21] = 0.0
    R.var("param")[92] = 0.0
    R.set("istep", 20), R.set("ifield", 1), R.set("ifstrs", 0)
    for key, tol, maxit in (("", 1e-8, 300), ("_15", -1e-30, 15)):
        o = [np.zeros(n) for _ in range(3)]
        ii = [a.copy() for a in rhs]
        R.call("ophinv", o[0], o[1], o[2], ii[0], ii[1], ii[2], h1, h2, tol, maxit)
        its = []
        for k, nm in enumerate(("VELX", "VELY", "VELZ")):
            x, r = np.zeros(n), rhs[k].copy()
            R.call("hmholtz", nm, x, r, h1, h2, masks[k], out["vmult"], 1, tol, maxit, k + 1)
            its.append(int(R.get("niterhm")))
            assert np.array_equal(x, o[k]) and np.array_equal(r, ii[k])
            out[f"o{k + 1}{key}"], out[f"r{k + 1}{key}"] = o[k], ii[k]
        out["its" + key] = np.array(its)
    # The same with h2 / 20 (the balance of a ten times larger time step): 133 / 128 / 153 iterations.  Over that many
    # iterations CG amplifies rounding-level differences to ~1e-6 of the residual, so a re-ordered summation may leave the loop
    # one iteration earlier or later; the reference's own log (residual per check) on the original and on a right-hand side
    # perturbed by one unit of rounding is stored so that the tests can tell such a flip from an error (count_or_margin).
    out["h2_long"] = h2 / 20.0
    for key, rr in (("long", rhs), ("long_pert", [ulp_perturbed(a, 100 + k) for k, a in enumerate(rhs)])):
        its = []
        for k, nm in enumerate(("VELX", "VELY", "VELZ")):
            x, r = np.zeros(n), rr[k].copy()
            with _Logged(R) as lg:
                R.call("hmholtz", nm, x, r, h1, out["h2_long"], masks[k], out["vmult"], 1, 1e-8, 300, k + 1)
            (hist, tol), = lg.cggo()
            its.append(int(R.get("niterhm")))
            out[f"rbn2_{key}{k + 1}"], out[f"tol_{key}{k + 1}"] = hist, np.array([tol])
            if key == "long":
                out[f"o{k + 1}_long"] = x
        out["its_" + key] = np.array(its)
    return out


def hsolve_inputs(case, pres=False, consistent=False):
    """A slowly varying sequence of right-hand sides (what successive time steps hand to hsolve) and the h1/h2 of each call;
    h2 changes at call 4, which makes project1 rebuild B = A X and re-orthogonalise (iproj_chk)."""
    n = case.n
    rng = np.random.default_rng(21)
    f = [case.bm1() * rng.standard_normal(n) for _ in range(3)]
    h1 = np.ones(n) if pres else 1.0 + 0.3 * rng.random(n)
    h2 = np.zeros(n) if pres else 20.0 * (5.0 + rng.random(n))
    calls = []
    for k in range(4 if pres else 11):     # 11 > mmx + 1: the space saturates at mmx = 8 and the oldest vector is rotated out
        rhs = f[0] + np.sin(0.3 * k) * f[1] + 0.05 * k * k * f[2]
        if consistent:       # all-Neumann pressure problem: the right-hand side integrates to zero, as crespsp's does after ortho
            rhs = rhs - case.bm1() * (rhs.sum() / case.bm1().sum())
        calls.append((rhs, h1, h2 * (1.1 if (k >= 4 and not pres) else 1.0), 10 + k))
    return calls


def _ref_hsolve(name, pres, case=None, tol=1e-7):
    case = case or case_of("core")
    rc = _ref(case)
    R, n = rc.R, case.n
    if pres:
        R.set("ifmgrid", 1)
        R.var("param")[[39, 40, 41, 42, 43]] = 0.0
        R.call("set_overlap")
        R.var("param")[20] = tol        # param(21)
        R.set("tolps", tol)
    mask = rc.fld("pmask" if pres else "v1mask")
    R.var("param")[21] = 0.0            # param(22)
    R.var("param")[92] = 20.0           # param(93): projection on, mxprev vectors
    R.var("param")[93] = 5.0            # param(94): from step 5 (velocity)
    R.var("param")[94] = 5.0            # param(95): from step 5 (pressure)
    R.var("ifprojfld")[1] = 1
    R.set("ifield", 1)
    approx, napprox = np.zeros(24 * n), np.zeros(10, dtype=np.int32)
    out = dict(mask=mask, vmult=rc.fld("vmult"), binvm1=rc.fld("binvm1"), volvm1=np.array([R.get("volvm1")]))
    its, ms = [], []
    for k, (rhs, h1, h2, istep) in enumerate(hsolve_inputs(case, pres, consistent=bool(R.get("ifvcor")) and pres)):
        R.set("istep", istep)
        u, r = np.zeros(n), rhs.copy()
        with _Logged(R) as lg:
            R.call("hsolve", name, u, r, h1, h2, mask, out["vmult"], 1, tol, 200, 1, approx, napprox, out["binvm1"])
        its.append(int(R.get("niterhm"))), ms.append(int(napprox[1]))
        out[f"u{k}"], out[f"r{k}"] = u, r
        # the residual the exit test saw at every check / iteration of this solve, and the tolerance it was held against
        (hist, tl), = (lg.gmres() if pres else lg.cggo())
        out[f"res{k}"], out[f"restol{k}"] = hist, np.array([tl])
    out["its"], out["m"] = np.array(its), np.array(ms)
    return out


def ref_hsolve():
    """core/navier4.f:562-634 hsolve with residual projection (project1/project2, :636-1199) around hmhzpf -> cggo (Jacobi
    PCG): eleven successive 'VELX' solves; the space grows to mmx = 8 vectors, is rebuilt when h2 changes (call 4) and then
    rotates its oldest vector out."""
    return _ref_hsolve("VELX", False)


def ref_hsolve_pres():
    """The same around the pressure solver of the Pn-Pn formulation: hsolve('PRES') -> project1 -> hmhzpf -> cggo('PRES') ->
    hmh_gmres (gmres.f:304-545) with h1mg_solve as preconditioner -> project2; four successive solves."""
    return _ref_hsolve("PRES", True)


def ref_hsolve_pres_channel():
    """BASELINE config 5 as examples/turbChannel/turbChannel.par runs it: Pn-Pn, [PRESSURE] residualTol = 1e-4 with
    residualProj = yes -- hsolve('PRES') with the residual projection around hmh_gmres / h1mg_solve, on the turbChannel mesh
    (4^3 cut: periodic x/z, stretched walls, constant null space); four successive solves."""
    out = _ref_hsolve("PRES", True, channel_case(), 1e-4)
    out["ifvcor"] = np.array([1])
    return out


MET9 = ("rxm2", "sxm2", "txm2", "rym2", "sym2", "tym2", "rzm2", "szm2", "tzm2")


def ref_eop():
    """The Pn-Pn-2 pressure operator: opgradt / cdtp, opdiv / multd, opbinv, cdabdtp(intype = 1) (core/navier1.f:258-850,
    4064-4114) with the mesh-2 geometry of geom2 (core/coef.f) on a deformed box."""
    case = case_of("eop")
    rc = _ref(case, lx2=6, ifsplit=False)
    R, E, n = rc.R, case.nel, case.n
    n2 = 216 * E
    f2 = lambda nm: R.var(nm)[..., :E].ravel(order="F").copy()
    out = dict(ixm12=R.var("ixm12").copy(), dxm12=R.var("dxm12").copy(), w3m2=R.var("w3m2").ravel(order="F").copy(),
               bm2=f2("bm2"), bm2inv=f2("bm2inv"), volvm2=np.array([R.get("volvm2")]), volvm1=np.array([R.get("volvm1")]),
               v1mask=rc.fld("v1mask"), v2mask=rc.fld("v2mask"), v3mask=rc.fld("v3mask"), vmult=rc.fld("vmult"),
               binvm1=rc.fld("binvm1"), bm1=rc.fld("bm1"))
    for nm in MET9:
        out[nm] = f2(nm)
    rng = np.random.default_rng(31)
    p = rng.standard_normal(n2)
    o = [np.zeros(n) for _ in range(3)]
    R.call("opgradt", o[0], o[1], o[2], p)
    u = [rng.standard_normal(n) for _ in range(3)]
    d = np.zeros(n2)
    R.call("opdiv", d, u[0], u[1], u[2])
    h2inv = 1.0 / (50.0 + rng.random(n))
    bo, bi = [np.zeros(n) for _ in range(3)], [a.copy() for a in u]
    R.call("opbinv", bo[0], bo[1], bo[2], bi[0], bi[1], bi[2], h2inv)
    ap = np.zeros(n2)
    R.call("cdabdtp", ap, p, np.ones(n), 1.0 / h2inv, h2inv, 1)
    # intype = -1: D (h1 A + h2 B)^-1 D^T, the velocity solves through ophinv with (tolhs, nmxv)
    R.set("tolhs", 1e-11), R.set("nmxv", 300), R.set("istep", 20), R.set("ifield", 1), R.set("ifstrs", 0)
    R.var("param")[21] = 0.0
    R.var("param")[92] = 0.0
    apm = np.zeros(n2)
    R.call("cdabdtp", apm, p, np.ones(n), 1.0 / h2inv, h2inv, -1)
    out["ap_m1"] = apm
    out.update(p=p, gx=o[0], gy=o[1], gz=o[2], ux=u[0], uy=u[1], uz=u[2], div=d, h2inv=h2inv, bo1=bo[0], bo2=bo[1], bo3=bo[2],
               bi1=bi[0], bi2=bi[1], bi3=bi[2], ap=ap)
    return out


def ref_uzawa():
    """core/gmres.f:2-237 uzawa_gmres: GMRES on E = cdabdtp(intype = 1) preconditioned by hsmg_solve (core/hsmg.f:1376-1602,
    set up by set_overlap -> gen_fast, hsmg_setup), tolerance through chktcg2 (core/navier1.f:1089-1154)."""
    case = case_of("eop")
    rc = _ref(case, lx2=6, ifsplit=False)
    R, E, n = rc.R, case.nel, case.n
    n2 = 216 * E
    R.set("ifmgrid", 1)
    R.var("param")[[39, 40, 41, 42, 43]] = 0.0
    R.call("set_overlap")
    rng = np.random.default_rng(41)
    h2inv = 1.0 / (50.0 + rng.random(n))
    h1, h2 = np.ones(n), 1.0 / h2inv
    pe = rng.standard_normal(n2)
    res = np.zeros(n2)
    R.call("cdabdtp", res, pe, h1, h2, h2inv, 1)
    out = dict(h2inv=h2inv, pe=pe, rhs=res.copy(), df=R.var("df")[:, :E].T.copy(), prelax=np.array([R.get("prelax")]),
               tolpdf=np.array([R.get("tolpdf")]))
    for nm in ("sr", "ss", "st"):
        out[nm] = R.var(nm)[:, :E].T.copy()
    R.var("param")[20] = 0.0
    R.set("tolps", 1e-7), R.set("istep", 5)
    it = C.c_int(0)
    x = res.copy()
    R.call("uzawa_gmres", x, h1, h2, h2inv, 1, it)
    out.update(x=x, it=np.array([it.value]))
    return out


MAP_NP = (1, 2, 3, 4, 5, 7, 8, 16, 48)


def ref_map():
    """core/map2.f:943-1026 assign_gllnid (incl. core/math.f isort / iswapt_ip for rank counts that are not a power of
    two) on the RSB leaves of the reference's examples/bp5/bp5.ma2 (tests/golden/bp5_fixture.npz)."""
    from oracle.ref import Ref
    R = Ref(8, 8, 64, fresh=True)
    leaf = np.load(os.path.join(os.path.dirname(GOLDEN), "bp5_fixture.npz"))["leaf"].astype(np.int32)
    out = {}
    for npr in MAP_NP:
        g, scratch = leaf.copy(), np.zeros(len(leaf), dtype=np.int32)
        R.call("assign_gllnid", g, scratch, len(g), len(g), npr)
        out[f"gllnid_np{npr}"] = g
    return out


REFERENCE = dict(channel=ref_channel, ethier=ref_ethier, core=ref_core, core_lx6=ref_core_lx6, h1mg_lx6=ref_h1mg_lx6, h1mg_lx4=ref_h1mg_lx4, h1mg_lx10=ref_h1mg_lx10, periodic=ref_periodic, map=ref_map, eop=ref_eop, uzawa=ref_uzawa, hsolve=ref_hsolve, hsolve_pres=ref_hsolve_pres, hsolve_pres_channel=ref_hsolve_pres_channel, ophinv=ref_ophinv, bp5=ref_bp5, h1mg=ref_h1mg, h1mg_neumann=ref_h1mg_neumann, fdm=ref_fdm, pnpn2=ref_pnpn2)


def reference_all():
    flat = {}
    for name, fn in REFERENCE.items():
        for k, v in fn().items():
            flat[f"{name}/{k}"] = np.asarray(v)
    return flat


def load_golden():
    z = np.load(GOLDEN)
    out = {}
    for k in z.files:
        name, key = k.split("/", 1)
        out.setdefault(name, {})[key] = z[k]
    return out


def fastd_to_S(g, E, nl=8):
    """common /fastd/ sr,ss,st(2*lx1*lx1,e) (S then S^T, column-major) -> S[e,3,row,col]; D = df."""
    S = np.zeros((E, 3, nl, nl))
    for d, nm in enumerate(("sr", "ss", "st")):
        S[:, d] = g[nm].reshape(E, 2, nl, nl)[:, 0].transpose(0, 2, 1)
    return S, g["df"].reshape(E, nl, nl, nl)
