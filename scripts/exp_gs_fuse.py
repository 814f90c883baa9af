"""Experiment: direct-stiffness summation gathered inside the CG update kernel (NEKB_GS_FUSE_UPDATE=1) against the stock
gs_local_kernel + cggos_update2_kernel pair.  Checks bit-identity of the two forms on small meshes (incl. ragged and periodic
numbering is covered by the handle itself), then times both at E = m^3.   python scripts/exp_gs_fuse.py [--m 64]"""
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
    ap.add_argument("--skip-small", action="store_true", help="timing only (for an ncu capture of the large case)")
    ap.add_argument("--modes", default="0,2,3,4", help="NEKB_GS_FUSE_UPDATE values to time (each twice)")
    a = ap.parse_args()
    from nek5000_b200 import lib, nek
    from nek5000_b200._lib import check
    from nek5000_b200.bp5 import BP5
    L = lib()
    out = {}
    # bit-identity on small boxes
    same = []
    for dims in () if a.skip_small else ((3, 2, 2), (4, 4, 3), (1, 1, 1), (2, 1, 1)):
        us = []
        for flag in ("0", "3", "4", "6"):
            os.environ["NEKB_GS_FUSE_UPDATE"] = flag
            nek.finalize()
            nek.init(0, 8, 3)
            b = BP5(*dims, lx1=8, deform=0.05)
            it, _, h = b.solve(-1e-8, 30, history=True)
            us.append((b.get("u1").copy(), h.copy()))
        same.append([bool(np.array_equal(us[0][0], u[0]) and np.array_equal(us[0][1], u[1])) for u in us[1:]] +
                    [float(np.abs(us[0][0] - u[0]).max() / np.abs(us[0][0]).max()) for u in us[1:]])
    out["bit_identical_small"] = same
    nek.finalize()
    nek.init(0, 8, 3)
    m = a.m
    b = BP5(m, m, m, lx1=8)
    res = {}
    for flag in a.modes.split(",") * 2:
        os.environ["NEKB_GS_FUSE_UPDATE"] = flag
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
        res.setdefault(flag, []).append({"ms_per_iteration": sec / it * 1e3, "gdofs": it * b.nel_global * 343 / sec / 1e9,
                                         "kernel_ms": prof, "relerr": b.relerr()})
    out["E"] = b.nel
    out["runs"] = res
    print(json.dumps(out))


if __name__ == "__main__":
    main()
