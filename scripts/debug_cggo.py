"""Debug helper (GPU box): per-iteration cggo history, CUDA path vs oracle."""
import ctypes as C
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oracle
from nek5000_b200 import nek, lib
from nek5000_b200.nek import DevArray
from nek5000_b200._lib import check

nek.init(0, 8, 3)
case = oracle.Case(3, 2, 2, nx=8, deform=0.05)
nek.set_nel(case.nel, case.nel); nek.set_gll(case.z, case.w); nek.set_dxyz(case.D, case.Dt)
nek.set_geom(*case.geom()[:7]); nek.set_ifdfrm(None)
h, _ = nek.setupds(8, case.nel, case.vertex)
nek.set_field_handle(1, h); nek.set_ifield(1)
bm1 = case.bm1(); nek.set_step_info(1, float(bm1.sum()))
rng = np.random.default_rng(2)
h1 = np.full(case.n, 1.3)
f = case.dssum(bm1 * rng.standard_normal(case.n)) * case.mask
for ifh2 in (False, True):
    h2 = np.full(case.n, 0.7) if ifh2 else np.zeros(case.n)
    for tin in (1e-6, 1e-10):
        xref, itref, hr = case.cggo(f, h1, h2, tin=tin, maxit=200, istep=1, history=True)
        d = [DevArray.from_host(a) for a in (np.zeros(case.n), f, h1, h2, case.mask, case.mult, case.binv())]
        hist = np.zeros(3 * 202); it = C.c_int(0)
        check(lib().nekb_cggo_dev(*[a.ptr for a in d], 1, tin, 200, C.byref(it), hist.ctypes.data))
        hg = hist.reshape(-1, 3)[:it.value + 1]
        m = min(len(hg), len(hr))
        rel = np.abs(hg[:m, 1] - hr[:m, 1]) / np.abs(hr[:m, 1])
        print(f"ifh2={ifh2} tin={tin}: it gpu {it.value} oracle {itref}; rbn2 rel diff max {rel.max():.2e} at {rel.argmax()}; last rbn2 gpu {hg[m-1,1]:.3e} oracle {hr[m-1,1]:.3e}")
        print("   rel diff every 10:", " ".join(f"{v:.1e}" for v in rel[::10]))
