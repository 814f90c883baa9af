"""Times the three velocity Helmholtz solves of ophinv (core/induct.f:1022-1090) on a synthetic box, device resident:
python scripts/bench_ophinv.py [--m 64] [--iters 60]

Runs the same fixed number of PCG iterations (tolh < 0 and tiny, so no component exits early) twice in separate processes:
  fused    -- hcg.cuh, one 3-right-hand-side PCG (default path)
  stock    -- NEKB_HCG=0: cggo_run, one component after the other, one kernel per reference statement
and prints one JSON line with ms per iteration-and-component for both, the speed-up, and the achieved HBM GB/s of the
fused path against its algorithmic 15.1 words per point, component and iteration (hcg.cuh header; 17.5 with
NEKB_HCG_RHO_KERNEL=1, the round-1 form with a pass of its own for rho).
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
WORDS_FUSED = 17.5 if os.environ.get("NEKB_HCG_RHO_KERNEL", "0") not in ("", "0") else 15.1
WORDS_STOCK = 30.0


def leg(m, iters):
    from nek5000_b200 import lib, nek
    from nek5000_b200._lib import check
    from nek5000_b200.bp5 import BP5
    L = lib()
    b = BP5(m, m, m, lx1=8)
    n = b.n
    nek.set_ifield(1)
    nek.set_field_handle(1, b.gs_handle)
    nek.set_step_info(20, 1.0)
    mask, mult = b.devptr("mask"), b.devptr("mult")
    rng = np.random.default_rng(0)
    h1 = nek.DevArray.from_host(1.0 + 0.3 * rng.random(n))
    h2 = nek.DevArray.from_host(5.0 + rng.random(n))
    binv = nek.DevArray.from_host(1.0 + rng.random(n))
    r1 = b.get("r1")
    rhs = [nek.DevArray.from_host(r1 * s) for s in (1.0, -0.5, 2.0)]
    out = [nek.DevArray(n) for _ in range(3)]
    it = np.zeros(3, dtype=np.int32)

    def run(k):
        check(L.nekb_ophinv_dev(out[0].ptr, out[1].ptr, out[2].ptr, rhs[0].ptr, rhs[1].ptr, rhs[2].ptr, h1.ptr, h2.ptr,
                                mask, mask, mask, mult, binv.ptr, -1e-200, k, it.ctypes.data, None))
        check(L.nekb_sync())
    run(4)
    t = []
    for _ in range(3):
        t0 = time.perf_counter()
        run(iters)
        t.append(time.perf_counter() - t0)
        assert it.tolist() == [iters] * 3, it
    # the set-up of a solve (dssum of the rhs, chktcg1, setprec, initial dots) is timed separately and subtracted
    t0 = time.perf_counter()
    run(1)
    t1 = time.perf_counter() - t0
    best = min(t)
    per = (best - t1) / (iters - 1) / 3
    # one right-hand side through cggo (the fused path with NRHS = 1, or cggo_run)
    it1 = C.c_int(0)

    def run1(k):
        check(L.nekb_cggo_dev(out[0].ptr, rhs[0].ptr, h1.ptr, h2.ptr, mask, mult, binv.ptr, 1, -1e-200, k, C.byref(it1), None))
        check(L.nekb_sync())
    run1(4)
    t0 = time.perf_counter()
    run1(iters)
    ta = time.perf_counter() - t0
    t0 = time.perf_counter()
    run1(1)
    tb = time.perf_counter() - t0
    per1 = (ta - tb) / (iters - 1)
    return dict(ms_per_iteration_component=per * 1e3, solve_ms=best * 1e3, setup_ms=t1 * 1e3, n=n, nel=b.nel,
                checksum=float(np.abs(out[2].to_host()).max()), cggo_1rhs_ms_per_iteration=per1 * 1e3)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--m", type=int, default=64)
    ap.add_argument("--iters", type=int, default=60)
    ap.add_argument("--leg", default=None)
    a = ap.parse_args()
    if a.leg:
        print(json.dumps(leg(a.m, a.iters)))
        return
    res = {}
    for name, env in (("fused", "1"), ("stock", "0")):
        e = dict(os.environ, NEKB_HCG=env)
        r = subprocess.run([sys.executable, os.path.abspath(__file__), "--m", str(a.m), "--iters", str(a.iters), "--leg", name],
                           env=e, capture_output=True, text=True)
        if r.returncode != 0:
            raise SystemExit(r.stderr[-2000:])
        res[name] = json.loads(r.stdout.strip().splitlines()[-1])
    peak = 6545.9
    try:
        peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        pass
    n = res["fused"]["n"]
    f, s = res["fused"]["ms_per_iteration_component"], res["stock"]["ms_per_iteration_component"]
    out = {"workload": f"ophinv, E={a.m}^3 elements, N=7, FP64, variable h1 and h2, {a.iters} fixed PCG iterations per component",
           "fused_ms_per_iteration_component": f, "stock_ms_per_iteration_component": s, "speedup": s / f,
           "fused_GBps": WORDS_FUSED * 8 * n / (f * 1e-3) / 1e9, "stock_GBps": WORDS_STOCK * 8 * n / (s * 1e-3) / 1e9,
           "hbm_peak_GBps": peak, "fused_frac_of_peak": WORDS_FUSED * 8 * n / (f * 1e-3) / 1e9 / peak,
           "cggo_1rhs_fused_ms_per_iteration": res["fused"]["cggo_1rhs_ms_per_iteration"],
           "cggo_1rhs_stock_ms_per_iteration": res["stock"]["cggo_1rhs_ms_per_iteration"],
           "cggo_1rhs_speedup": res["stock"]["cggo_1rhs_ms_per_iteration"] / res["fused"]["cggo_1rhs_ms_per_iteration"],
           "relative_difference_of_results": abs(res["fused"]["checksum"] - res["stock"]["checksum"]) / res["stock"]["checksum"],
           "legs": res}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
