"""The drop-in, proven through the REFERENCE'S OWN CALL SITES.

oracle/_ref/libnekref_lx8e64hyb.so (oracle/ref_build.py --hybrid) is the reference's Fortran, transpiled as for the all-
reference library, EXCEPT that axhelm, dssum, dsop, cggo, cggos, axhm1, h1mg_solve, h1mg_setup and gslib's / crs' Fortran
API are left out: those names are undefined in it and the dynamic linker binds them to nek5000_b200/libnekb200.so -- what
a Nek5000 build gets when libnekb200.so precedes libnek5000.a on the link line.  The reference's own

  * hmholtz (core/hmholtz.f:2-69: dssum, col2, chktcg1 -> axhelm, cggo),
  * bp5 driver (examples/bp5/bp5.usr:324-395: geodatstd, rand_fld_h1, xmask1, axhm1, dssum, cggos, glrdif),
  * set_overlap -> h1mg_setup (overridden by the glue) and hmh_gmres (core/gmres.f:304-545: ax -> axhelm + dssum,
    h1mg_solve, ortho, the Givens recurrences)

then call the CUDA entry points with the Fortran calling convention (everything by reference, hidden CHARACTER lengths,
operands in COMMON registered once by oracle/hyb_glue.c, the C rendering of INTEGRATION.md's glue).  The results must
reproduce the ALL-REFERENCE goldens (tests/golden/ref_golden.npz): identical iteration counts, fields to 1e-10."""
import ctypes as C

import numpy as np
import pytest

import refcases

pytestmark = pytest.mark.gpu

G = refcases.load_golden()
TOL_FIELD = 1e-10
TOL_CONVERGED = 1e-7


def relmax(a, b):
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-300)


@pytest.fixture()
def hyb():
    """Fresh product state + a private copy of the hybrid library per test."""
    from nek5000_b200 import nek
    from oracle.ref import RefCase
    nek.finalize()

    def make(case, **kw):
        return RefCase(case, hybrid=True, **kw)
    yield make
    nek.finalize()


def test_the_replaced_routines_are_undefined_in_the_hybrid_library():
    """nm-level check of the claim above (runs without touching the GPU)."""
    import subprocess
    from oracle import ref_build
    so = ref_build.build(8, 8, 64, hybrid=True)
    out = subprocess.run(["nm", "-D", so], capture_output=True, text=True).stdout
    und = {l.split()[-1] for l in out.splitlines() if " U " in l}
    for name in ("axhelm_", "dssum_", "dsop_", "cggo_", "cggos_", "axhm1_", "h1mg_solve_", "fgslib_gs_setup_", "fgslib_gs_op_"):
        assert name in und, name
    defined = {l.split()[-1] for l in out.splitlines() if " T " in l}
    for name in ("hmholtz_", "chktcg1_", "hmh_gmres_", "bp5_", "set_overlap_", "setupds_", "setvert3d_"):
        assert name in defined, name
    dyn = subprocess.run(["readelf", "-d", so], capture_output=True, text=True).stdout
    assert "libnekb200.so" in dyn


@pytest.mark.parametrize("name", ["ethier", "channel"])
def test_reference_hmholtz_on_cuda_axhelm_dssum_cggo(hyb, name):
    """The reference's hmholtz('VELX') of a time step (h1 = viscosity, h2 = bd/dt): set-up (setupds -> fgslib_gs_setup,
    dssum of the multiplicity, setinvm's dssum) already runs on the product's gather-scatter."""
    g = G[name]
    case = refcases.channel_case() if name == "channel" else refcases.ethier_case()
    rc = hyb(case)
    R, n = rc.R, case.n
    # the geometry the reference computed on top of the product's gs must be the all-reference geometry
    assert np.array_equal(rc.fld("vmult"), g["vmult"]) and relmax(rc.fld("binvm1"), g["binvm1"]) <= 1e-14
    x, r = np.zeros(n), g["vel_rhs"].copy()
    R.var("param")[21] = 0.0
    R.set("ifsolv", 0), R.set("kfldfdm", -1), R.set("istep", 1), R.set("ifield", 1)
    R.call("nekhyb_step")
    R.call("hmholtz", "VELX", x, r, np.full(n, g["vel_h1"][0]), np.full(n, g["vel_h2"][0]), rc.fld("v1mask"), rc.fld("vmult"),
           1, 1e-9, 200, 1)
    R.call("nekhyb_fetch_niterhm")
    assert int(R.get("niterhm")) == int(g["vel_it"][0])                      # identical iteration count
    assert relmax(x, g["vel_x"]) <= TOL_CONVERGED


def test_reference_bp5_driver_on_cuda_axhm1_dssum_cggos(hyb):
    """examples/bp5/bp5.usr:324-395 unchanged: 40 fixed CG iterations through cggos_ (fused Ax + vector updates), the
    right-hand side through axhm1_ + dssum_."""
    g, case = G["bp5"], refcases.case_of("neumann")
    rc = hyb(case)
    R, n = rc.R, case.n
    R.var("uparam")[0:3] = (-1e-8, 40, 1)
    gf = R.var("gf", "bp5")
    R.call("nekhyb_bp5_register", gf)
    R.call("bp5")
    v = lambda nm: R.var(nm, "bp5").ravel(order="F")
    assert np.array_equal(v("gf")[:6 * n], g["gf"])                          # geodatstd is the reference's own
    assert np.array_equal(v("e1")[:n], g["e1"])                              # rand_fld_h1 + dsavg on the product's gs
    assert relmax(v("r1")[:n], g["r1"]) <= 1e-12                             # axhm1_ + dssum_ per apply
    assert relmax(v("u1")[:n], g["u1"]) <= TOL_FIELD                         # 40 iterations of cggos_


@pytest.mark.parametrize("name", ["ethier", "channel"])
def test_reference_hmh_gmres_on_cuda_h1mg_solve_axhelm_dssum(hyb, name):
    """The reference's GMRES (its own Givens recurrences, ortho, chktcg-free exit test) preconditioned by h1mg_solve_ and
    applying ax -> axhelm_ + dssum_: identical iteration count and solution, constant null space included."""
    g = G[name]
    case = refcases.channel_case() if name == "channel" else refcases.ethier_case()
    rc = hyb(case)
    R, n = rc.R, case.n
    R.set("ifmgrid", 1)
    R.var("param")[[39, 40, 41, 42, 43]] = 0.0
    tol = float(g["tol"][0])
    R.var("param")[20] = tol
    R.set("tolps", tol), R.set("istep", 1)
    R.call("nekhyb_step")
    R.call("set_overlap")                                                      # -> the glue's h1mg_setup -> nekb_h1mg_setup
    z, r = np.zeros(n), g["rhs"].copy()
    R.call("h1mg_solve", z, r, False)
    assert np.array_equal(r, g["rhs_out"]) and relmax(z, g["z"]) <= TOL_FIELD
    x, it = g["b"].copy(), C.c_int(100)
    R.call("hmh_gmres", x, np.ones(n), np.zeros(n), rc.fld("vmult"), it)
    assert it.value == int(g["it"][0])
    assert relmax(x, g["x"]) <= 1e-9
