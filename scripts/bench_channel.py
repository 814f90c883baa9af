"""BASELINE config 5 measurement: the pressure solve on the complete mesh of examples/turbChannel
(16 x 12 x 8 = 1536 elements, periodic x/z, stretched walls, lx1 = 8) -- time per h1mg_solve call and per hmh_gmres /
hmh_flex_cg solve, with the iteration counts, through the Fortran-named entry points (host buffers: every call includes
its host<->device copies).  `--reference` adds the reference's own Fortran (oracle/_ref, lelt = 1536 build) timed on one
host core for the same three calls.

    python scripts/bench_channel.py [--calls 20] [--reference]
Prints one JSON line.
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
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--calls", type=int, default=20)
    ap.add_argument("--reference", action="store_true")
    a = ap.parse_args()
    import refcases
    from nek5000_b200 import nek
    case = refcases.channel_case(refcases.CHANNEL_FULL_DIMS)
    E, n = case.nel, case.n
    geo = case.geom()
    pmask = np.ones(n)
    rhs, b = refcases.pressure_inputs(case, pmask)
    out = {"workload": f"turbChannel mesh {refcases.CHANNEL_FULL_DIMS} = {E} elements, lx1=8, {n} points, tol 1e-8, null space"}

    nek.finalize()
    nek.init(0, 8, 3)
    nek.set_nel(E, E)
    nek.set_gll(case.z, case.w)
    nek.set_dxyz(case.D, np.ascontiguousarray(case.D.T))
    nek.set_geom(*geo[:6], geo[6])
    nek.set_ifdfrm(None)
    h, _ = nek.setupds(8, E, case.vertex)
    nek.set_ifield(1)
    nek.set_field_handle(1, h)
    nek.set_step_info(1, float(geo[6].sum()))
    nek.set_binv(case.binv())
    t0 = time.perf_counter()
    nek.h1mg_setup(refcases.channel_fbc(case), case.xm1, case.ym1, case.zm1, case.vertex, E, True)
    out["h1mg_setup_s"] = time.perf_counter() - t0
    nek.set_pressure_state(pmask, case.binv(), 1e-8, 1e-8, True, E)
    z = np.zeros(n)
    for _ in range(3):
        nek.h1mg_solve(z, rhs.copy(), False)
    rs = [rhs.copy() for _ in range(a.calls)]
    nek.launch_count(reset=True)
    t0 = time.perf_counter()
    for r in rs:
        nek.h1mg_solve(z, r, False)
    out["h1mg_solve_ms"] = (time.perf_counter() - t0) / a.calls * 1e3
    out["h1mg_solve_launches"] = nek.launch_count() / a.calls
    one, zero = np.ones(n), np.zeros(n)
    for name, fn in (("hmh_gmres", nek.hmh_gmres), ("hmh_flex_cg", nek.hmh_flex_cg)):
        fn(b.copy(), one, zero, case.mult, 100)                         # allocates the bases
        best, it = 1e30, 0
        for _ in range(3):
            res = b.copy()
            t0 = time.perf_counter()
            it = fn(res, one, zero, case.mult, 100)
            best = min(best, time.perf_counter() - t0)
        out[name] = {"iterations": it, "ms": best * 1e3, "ms_per_iteration": best * 1e3 / max(it, 1)}
    nek.finalize()

    if a.reference:
        from oracle.ref import RefCase
        rc = RefCase(case)
        R = rc.R
        R.set("ifmgrid", 1)
        R.var("param")[[39, 40, 41, 42, 43]] = 0.0
        R.call("set_overlap")
        R.var("param")[20] = 1e-8
        R.set("tolps", 1e-8), R.set("istep", 1)
        t0 = time.perf_counter()
        for _ in range(3):
            R.call("h1mg_solve", np.zeros(n), rhs.copy(), False)
        ref = {"cores": 1, "h1mg_solve_ms": (time.perf_counter() - t0) / 3 * 1e3}
        for name in ("hmh_gmres", "hmh_flex_cg"):
            x, it = b.copy(), C.c_int(100)
            t0 = time.perf_counter()
            R.call(name, x, one, zero, case.mult, it)
            sec = time.perf_counter() - t0
            ref[name] = {"iterations": it.value, "ms": sec * 1e3, "ms_per_iteration": sec * 1e3 / max(it.value, 1)}
        out["reference_cpu"] = ref
    print(json.dumps(out))


if __name__ == "__main__":
    main()
