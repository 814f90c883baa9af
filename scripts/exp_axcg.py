"""Experiment: one operator-kernel variant (NEKB_AXCG_VARIANT, read once per process) -- parity on a small affine brick
against the oracle and against the general kernel, then the BP5 iteration timed at E = m^3 with per-kernel CUDA events.
    NEKB_AXCG_VARIANT=10 python scripts/exp_axcg.py --m 64 --its 100"""
import argparse
import ctypes as C
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--m", type=int, default=64)
    ap.add_argument("--its", type=int, default=100)
    ap.add_argument("--skip-small", action="store_true")
    ap.add_argument("--general", action="store_true", help="time the general-geometry kernel (NEKB_AX_AFFINE=0)")
    a = ap.parse_args()
    import oracle
    from nek5000_b200 import lib, nek
    from nek5000_b200._lib import check
    from nek5000_b200.bp5 import BP5
    L = lib()
    out = {"variant": os.environ.get("NEKB_AXCG_VARIANT", "0")}
    if not a.skip_small:
        small = []
        for dims in ((3, 3, 2), (5, 4, 3), (1, 1, 1)):
            nek.finalize()
            nek.init(0, 8, 3)
            case = oracle.Case(*dims, nx=8)
            nek.set_gll(case.z, case.w)
            nek.set_dxyz(case.D, case.Dt)
            b = BP5(*dims, lx1=8)
            e1, r1 = case.bp5_problem()
            uref, itref, hist = case.cggos(r1, e1, maxit=30, history=True)
            us = []
            for flag in ("1", "0"):
                os.environ["NEKB_AX_AFFINE"] = flag
                it, sec, h = b.solve(-1e-8, 30, history=True)
                us.append((b.get("u1").copy(), np.array(h).copy()))
            del os.environ["NEKB_AX_AFFINE"]
            sc = np.abs(uref).max()
            small.append({"dims": dims, "affine_active": int(L.nekb_ax_affine_active()),
                          "affine_vs_general": float(np.abs(us[0][0] - us[1][0]).max() / sc),
                          "affine_vs_oracle": float(np.abs(us[0][0] - uref).max() / sc),
                          "general_vs_oracle": float(np.abs(us[1][0] - uref).max() / sc),
                          "hist_affine_vs_general": float(np.abs(us[0][1] - us[1][1]).max() / np.abs(us[1][1]).max())})
        out["small"] = small
    nek.finalize()
    nek.init(0, 8, 3)
    m = a.m
    b = BP5(m, m, m, lx1=8)
    if a.general:
        os.environ["NEKB_AX_AFFINE"] = "0"
    out["general"] = bool(a.general)
    runs = []
    for rep in range(2):
        b.solve(-1e-8, 5)
        check(L.nekb_prof_enable(1))
        it, sec = b.solve(-1e-8, a.its)
        prof = {}
        for k in ("ax", "gs", "update", "pupdate"):
            s_, c_ = C.c_double(0), C.c_longlong(0)
            check(L.nekb_prof_get(k.encode(), C.byref(s_), C.byref(c_)))
            prof[k] = s_.value / max(c_.value, 1) * 1e3
        check(L.nekb_prof_enable(0))
        it, sec = b.solve(-1e-8, a.its)
        runs.append({"ms_per_iteration": sec / it * 1e3, "gdofs": it * b.nel_global * 343 / sec / 1e9, "kernel_ms": prof,
                     "relerr": b.relerr()})
    out["E"] = b.nel
    out["runs"] = runs
    print(json.dumps(out))


if __name__ == "__main__":
    main()
