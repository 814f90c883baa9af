"""BP5 driver: the benchmark loop of examples/bp5/bp5.usr:324-395 on the device-resident synthetic case.

    case = BP5(nelx, nely, nelz, lx1=8)          # genbox box [0,1]^3, 'v  ' on all sides, x-fastest elements
    res  = case.run(maxit=500, ntests=100)       # cggos x ntests, reference DoF/s accounting (bp5.usr:378-383)

Everything (coordinates, geodatstd, numbering -> gs_setup, mask, multiplicity, ran1 exact solution, rhs, CG) is
built and solved by libnekb200.so; with np > 1 ranks the box is split into px*py*pz bricks (RCB-equivalent of the
genmap/parRSB partition of a box) with NCCL for the shared-node exchange and the CG all-reduces.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import nek
from ._lib import check, lib


def brick_layout(nranks: int):
    """px,py,pz for a power-of-two rank count: recursive bisection z, y, x, z, ... (2x2x2 at 8 ranks)."""
    p = [1, 1, 1]
    d, r = 2, nranks
    while r > 1:
        if r % 2:
            raise ValueError("rank count must be a power of two")
        p[d] *= 2
        r //= 2
        d = (d - 1) % 3
    return tuple(p)


class BP5:
    def __init__(self, nelx: int, nely: int, nelz: int, lx1: int = 8, device: int = 0, rank: int = 0, nranks: int = 1,
                 layout=None, deform: float = 0.0):
        self.lx1, self.nranks, self.rank = lx1, nranks, rank
        self.nel_global = nelx * nely * nelz
        nek.init(device, lx1, 3)
        px, py, pz = layout or brick_layout(nranks)
        check(lib().nekb_bp5_setup(nelx, nely, nelz, px, py, pz, float(deform)))
        self.nel = int(lib().nekb_bp5_nel_local())
        self.n = self.nel * lx1 ** 3
        nek._state["nelt"] = self.nel

    # host copies (parity tests) ------------------------------------------------------------------------------
    def get(self, which: str) -> np.ndarray:
        if which == "glo_num":
            out = np.zeros(self.n, dtype=np.int64)
        elif which in ("gf", "g"):
            out = np.zeros(6 * self.n)
        else:
            out = np.zeros(self.n)
        check(lib().nekb_bp5_get(which.encode(), out.ctypes.data, out.nbytes))
        return out

    def devptr(self, which: str) -> int:
        p = lib().nekb_bp5_devptr(which.encode())
        if not p:
            raise nek.NekbError(lib().nekb_last_error().decode())
        return p

    @property
    def gs_handle(self) -> int:
        return int(lib().nekb_bp5_gs_handle())

    # solver -------------------------------------------------------------------------------------------------------
    def solve(self, tol: float = -1e-8, maxit: int = 500, history: bool = False):
        """One cggos call (bp5.usr:367-369 body).  Returns (iterations, seconds[, hist (iters,3): pap, rtz, err])."""
        it, sec = C.c_int(0), C.c_double(0.0)
        hist = np.zeros(3 * (maxit + 1)) if history else None
        check(lib().nekb_bp5_solve(tol, maxit, C.byref(it), C.byref(sec), None if hist is None else hist.ctypes.data))
        if history:
            return int(it.value), float(sec.value), hist[:3 * it.value].reshape(-1, 3)
        return int(it.value), float(sec.value)

    def relerr(self) -> float:
        r = C.c_double(0.0)
        check(lib().nekb_bp5_relerr(C.byref(r)))
        return float(r.value)

    def run(self, maxit: int = 500, ntests: int = 100, tol: float = -1e-8, log=print):
        """bp5.usr:362-391: ntests solves, DoF/s = ntests*niter * nelgt*N^3 / elapsed."""
        niter, elapsed = 0, 0.0
        for _ in range(ntests):
            it, sec = self.solve(tol, maxit)
            niter += it
            elapsed += sec
        N = self.lx1 - 1
        dof = self.nel_global * N ** 3
        rel = self.relerr()
        rate = niter * dof / elapsed if elapsed > 0 else float("nan")
        if log and self.rank == 0:
            log(f"nproc | N | DoF | niter | relerr | elapsed | DoF/s\n"
                f"{self.nranks} {N} {dof} {niter} {rel:.6e} {elapsed:.6e} {rate:.6e}")
        return {"nproc": self.nranks, "N": N, "dof": dof, "niter": niter, "relerr": rel, "elapsed": elapsed, "dofs": rate}
