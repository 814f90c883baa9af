"""Times the pressure preconditioner (h1mg_solve) and a preconditioned GMRES solve on a synthetic box
(SURVEY.md 8d, config-5 style measurement): python scripts/bench_hsmg.py [--m 48] [--calls 20]

Box [0,1]^3 of m^3 elements, N=7; pressure boundary conditions: outflow ('O', Dirichlet for p) on x+, walls elsewhere.
Prints one JSON line: ms per h1mg_solve call, GMRES iterations / seconds for a 1e-6 relative reduction, coarse-solve
iterations, and the algorithmic HBM bytes of one V-cycle for orientation.
"""
import argparse
import ctypes as C
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--m", type=int, default=48)
    ap.add_argument("--calls", type=int, default=20)
    ap.add_argument("--no-gmres", action="store_true")
    a = ap.parse_args()
    from nek5000_b200 import lib, nek
    from nek5000_b200._lib import check
    from nek5000_b200.bp5 import BP5
    L = lib()
    m, lx1 = a.m, 8
    t0 = time.perf_counter()
    b = BP5(m, m, m, lx1=lx1)
    nel, n = b.nel, b.n
    x, y, z = b.get("xm1"), b.get("ym1"), b.get("zm1")
    e = np.arange(nel)
    ex, ey, ez = e % m, (e // m) % m, e // (m * m)
    q = np.arange(8)
    vertex = (1 + (ex[:, None] + (q & 1)) + (m + 1) * ((ey[:, None] + ((q >> 1) & 1)) + (m + 1) * (ez[:, None] + (q >> 2)))).astype(np.int64)
    nek.set_geom_from_xyz(x, y, z, bp5_form=False)
    fbc = np.zeros((nel, 6), dtype=np.int32)
    fbc[ex == 0, 0] = 2
    fbc[ex == m - 1, 1] = 1
    fbc[ey == 0, 2] = 2
    fbc[ey == m - 1, 3] = 2
    fbc[ez == 0, 4] = 2
    fbc[ez == m - 1, 5] = 2
    t1 = time.perf_counter()
    nek.h1mg_setup(fbc, x, y, z, vertex, nel, False)
    t2 = time.perf_counter()
    info = nek.h1mg_info()
    rng = np.random.default_rng(0)
    mult = b.get("mult")
    pm = np.ones((m, m, m, lx1, lx1, lx1))
    pm[:, :, -1, :, :, -1] = 0
    pmask = pm.reshape(-1)
    h = b.gs_handle
    rd = nek.DevArray.from_host(rng.standard_normal(n))
    check(L.nekb_gs_op_dev(h, rd.ptr, 1, None))
    r0 = rd.to_host() * mult * pmask
    zd = nek.DevArray(n)
    for _ in range(3):
        check(L.nekb_h2d(rd.ptr, r0.ctypes.data, r0.nbytes))
        check(L.nekb_h1mg_solve_dev(zd.ptr, rd.ptr))
    check(L.nekb_sync())
    nek.launch_count(reset=True)
    ts = time.perf_counter()
    for _ in range(a.calls):
        check(L.nekb_h1mg_solve_dev(zd.ptr, rd.ptr))
    check(L.nekb_sync())
    ms_call = (time.perf_counter() - ts) / a.calls * 1e3
    launches = nek.launch_count() / a.calls
    crs_it = nek.h1mg_info()["crs_iters"]
    out = {"workload": f"h1mg_solve, box {m}^3 = {nel} elements, N=7, levels nh={info['nh']}, FDM table rows {info['ntab']}",
           "ms_per_h1mg_solve": ms_call, "launches_per_call": launches, "coarse_pcg_iterations": crs_it,
           "setup_s": t2 - t1, "case_s": t1 - t0}
    # one V-cycle, algorithmic words per fine grid point (fine level only dominates): mask+faces 2, FDM 2, overlap add 2,
    # gs+weights ~3.4, restriction read 2, prolongation 2, dsavg ~3.4
    out["ms_per_ax_for_scale"] = None
    if not a.no_gmres:
        xe = rng.standard_normal(n)
        xd = nek.DevArray.from_host(xe)
        check(L.nekb_gs_op_dev(h, xd.ptr, 1, None))
        xe = xd.to_host() * mult * pmask
        check(L.nekb_h2d(xd.ptr, xe.ctypes.data, xe.nbytes))
        pmd, wtd = nek.DevArray.from_host(pmask), nek.DevArray.from_host(mult)
        bd = nek.DevArray(n)
        check(L.nekb_ax_bp5_dev(bd.ptr, xd.ptr, None))
        check(L.nekb_gs_op_dev(h, bd.ptr, 1, pmd.ptr))
        nek.set_step_info(1, 1.0)
        it = C.c_int(0)
        hist = np.zeros(201)
        b0 = bd.to_host()
        check(L.nekb_hmh_gmres_dev(bd.ptr, None, None, wtd.ptr, pmd.ptr, -1e-6, 200, C.byref(it), hist.ctypes.data, None))  # allocates the bases
        check(L.nekb_h2d(bd.ptr, b0.ctypes.data, b0.nbytes))
        check(L.nekb_sync())
        ts = time.perf_counter()
        check(L.nekb_hmh_gmres_dev(bd.ptr, None, None, wtd.ptr, pmd.ptr, -1e-6, 200, C.byref(it), hist.ctypes.data, None))
        check(L.nekb_sync())
        sec = time.perf_counter() - ts
        xs = bd.to_host()
        out.update({"gmres_iterations": it.value, "gmres_seconds": sec, "gmres_ms_per_iteration": sec / max(it.value, 1) * 1e3,
                    "gmres_rel_error": float(np.abs(xs - xe).max() / np.abs(xe).max()),
                    "gmres_residual_first_last": [float(hist[0]), float(hist[max(it.value - 1, 0)])]})
    print(json.dumps(out))


if __name__ == "__main__":
    main()
