"""torchrun worker: h1mg_solve and hmh_gmres on N GPUs (element-partitioned, NCCL) against the undivided numpy oracle.
Prints 'MGPU-HSMG-OK rank r'."""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist
    import oracle
    from oracle import hsmg
    from nek5000_b200 import nek
    from nek5000_b200._lib import check, lib
    from nek5000_b200.bp5 import brick_layout

    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    nek.init(local, 8, 3)
    nek.comm_init_torch()
    px, py, pz = brick_layout(world)
    lx, ly, lz = 2, 2, 1
    nelx, nely, nelz = lx * px, ly * py, max(lz * pz, 2)
    lz = nelz // pz
    case = oracle.Case(nelx, nely, nelz, nx=8, dirichlet=(0, 1, 0, 0, 0, 0), deform=0.03)
    fbc = hsmg.box_fbc(case, (2, 1, 2, 2, 2, 2))
    mg = hsmg.H1MG(case, fbc)
    nxyz = 512
    eg = np.arange(case.nel)
    ex, ey, ez = eg % nelx, (eg // nelx) % nely, eg // (nelx * nely)
    owner = (ex // lx) + px * ((ey // ly) + py * (ez // lz))
    local_of = (ex % lx) + lx * ((ey % ly) + ly * (ez % lz))
    mine = np.flatnonzero(owner == rank)
    order = mine[np.argsort(local_of[mine])]
    take = (order[:, None] * nxyz + np.arange(nxyz)[None, :]).reshape(-1)
    nel = len(order)
    loc = lambda a: np.ascontiguousarray(a[take])
    nek.set_nel(nel, nel)
    nek.set_gll(case.z, case.w)
    nek.set_dxyz(case.D, case.Dt)
    nek.set_geom(*[loc(g) for g in case.geom()[:7]])
    nek.set_ifdfrm(None)
    vertex = np.ascontiguousarray(case.vertex.reshape(-1, 8)[order].reshape(-1))
    nek.h1mg_setup(np.ascontiguousarray(fbc[order]), loc(case.xm1), loc(case.ym1), loc(case.zm1), vertex, nel, False)
    rel = lambda a, c: np.abs(a - c).max() / max(np.abs(c).max(), 1e-300)

    rng = np.random.default_rng(21)
    rhs = case.dssum(rng.standard_normal(case.n)) * case.mult
    zref = mg.solve(rhs.copy())
    z, r = np.zeros(nel * nxyz), loc(rhs)
    nek.h1mg_solve(z, r, False)
    assert rel(z, zref[take]) <= 1e-10, ("h1mg_solve", rel(z, zref[take]))

    n = case.n
    h1, h2 = np.ones(n), np.zeros(n)
    xe = case.dssum(rng.standard_normal(n)) * case.mult * case.mask
    b = case.dssum(case.axhelm(xe, h1, h2)) * case.mask
    nek.set_step_info(1, float(case.bm1().sum()))
    tol, maxit = 1e-8, 60
    xref, itref, hist_ref, div0 = hsmg.hmh_gmres(case, mg, b, h1, h2, case.mask, case.mult, tol, maxit, history=True)
    L = lib()
    bd, h1d, wtd, pmd = (nek.DevArray.from_host(loc(a)) for a in (b, h1, case.mult, case.mask))
    it = C.c_int(0)
    hist = np.zeros(maxit + 1)
    check(L.nekb_hmh_gmres_dev(bd.ptr, h1d.ptr, None, wtd.ptr, pmd.ptr, tol, maxit, C.byref(it), hist.ctypes.data, None))
    x = bd.to_host()
    assert it.value == itref, (it.value, itref)
    assert np.abs(hist[:itref] - hist_ref).max() <= 1e-9 * div0
    assert rel(x, xref[take]) <= 1e-9
    print(f"MGPU-HSMG-OK rank {rank} of {world}: gmres its={it.value} rel(z)={rel(z, zref[take]):.2e}", flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
