"""Times the Pn-Pn-2 pressure-operator kernels (opgradt = D^T, opdiv = D, cdabdtp = D (h2 B)^-1 D^T) on a synthetic box:
python scripts/bench_pnpn2.py [--m 32] [--calls 20]

Geometry enters these kernels only through the nine mesh-2 metric arrays, so the timing uses synthetic metrics of the right
size (the parity tests use the reference's own geometry).  Prints one JSON line: ms per call and the achieved HBM GB/s
against the algorithmic traffic (opgradt: 216*10 words in, 512*3 out per element; opdiv: the transpose).
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def lagrange(zf, zc):
    """I[a, i] = l_i(zf_a), D[a, i] = l_i'(zf_a) for nodes zc."""
    n = len(zc)
    I, D = np.zeros((len(zf), n)), np.zeros((len(zf), n))
    for i in range(n):
        others = [zc[k] for k in range(n) if k != i]
        den = np.prod([zc[i] - o for o in others])
        for a, x in enumerate(zf):
            I[a, i] = np.prod([x - o for o in others]) / den
            D[a, i] = sum(np.prod([x - o for q, o in enumerate(others) if q != p]) for p in range(n - 1)) / den
    return I, D


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--m", type=int, default=32)
    ap.add_argument("--calls", type=int, default=20)
    a = ap.parse_args()
    from nek5000_b200 import lib, nek
    from nek5000_b200._lib import check
    from nek5000_b200.bp5 import BP5
    import oracle
    L = lib()
    b = BP5(a.m, a.m, a.m, lx1=8)
    E, n = b.nel, b.n
    n2 = 216 * E
    z1, _ = oracle.zwgll(8)
    z2, w2 = np.polynomial.legendre.leggauss(6)
    I, D = lagrange(z2, z1)
    w3 = np.einsum("i,j,k->ijk", w2, w2, w2).reshape(-1)
    rng = np.random.default_rng(0)
    mets = [rng.standard_normal(n2) * 0.05 for _ in range(9)]
    nek.set_ifield(1)
    nek.set_field_handle(1, b.gs_handle)
    nek.set_step_info(5, 1.0)
    mask, mult = b.get("mask"), b.get("mult")
    nek.set_velocity_state(mask, mask, mask, mult)
    nek.set_mesh2(6, I, D, w3, mets, np.ones(n2), np.ones(n2), 1.0, 1e-8, 100, E, False)
    Dv = nek.DevArray
    p, ap_ = Dv.from_host(rng.standard_normal(n2)), Dv(n2)
    u = [Dv.from_host(rng.standard_normal(n)) for _ in range(3)]
    o = [Dv(n) for _ in range(3)]
    h1, h2, h2inv = Dv.from_host(np.ones(n)), Dv.from_host(np.full(n, 50.0)), Dv.from_host(np.full(n, 0.02))

    def timeit(fn):
        fn()
        check(L.nekb_sync())
        t0 = time.perf_counter()
        for _ in range(a.calls):
            fn()
        check(L.nekb_sync())
        return (time.perf_counter() - t0) / a.calls

    t_g = timeit(lambda: check(L.nekb_opgradt_dev(o[0].ptr, o[1].ptr, o[2].ptr, p.ptr)))
    t_d = timeit(lambda: check(L.nekb_opdiv_dev(ap_.ptr, u[0].ptr, u[1].ptr, u[2].ptr)))
    t_e = timeit(lambda: check(L.nekb_cdabdtp_dev(ap_.ptr, p.ptr, h1.ptr, h2.ptr, h2inv.ptr, 1)))
    words = 216 * 10 + 512 * 3
    peak = 6545.9
    try:
        peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        pass
    gb = words * 8 * E / 1e9
    print(json.dumps({"workload": f"Pn-Pn-2 operators, E={a.m}^3={E} elements, lx1=8, lx2=6, FP64",
                      "opgradt_ms": t_g * 1e3, "opdiv_ms": t_d * 1e3, "cdabdtp_intype1_ms": t_e * 1e3,
                      "algorithmic_GB_per_call": gb, "opgradt_GBps": gb / t_g, "opdiv_GBps": gb / t_d,
                      "opgradt_frac_of_peak": gb / t_g / peak, "opdiv_frac_of_peak": gb / t_d / peak, "hbm_peak_GBps": peak}))


if __name__ == "__main__":
    main()
