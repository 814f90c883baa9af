"""CPU oracle for the Nek5000 BP5 / Helmholtz-PCG-dssum hot path (TEST INFRASTRUCTURE ONLY).

numpy/ctypes front end of ``oracle/nek_oracle.c`` (statement-level C restatement of
the reference Fortran; each C function cites the reference file:line it follows) and
of ``oracle/bp5_cpu.c`` (OpenMP "restated CPU baseline" for bench.py).

PARITY PINNED against the reference itself: ``oracle/ref_build.py`` transpiles the reference's own
Fortran (``oracle/f77c.py``; there is no Fortran compiler here) into ``oracle/_ref/`` and
``tests/test_ref_pins.py`` holds this restatement to its outputs bit for bit (golden vectors:
``tests/golden/ref_golden.npz``); see the header of nek_oracle.c.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline /
``--impl reference`` legs may import this package.  The product (``nek5000_b200``)
never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import tempfile

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_BUILD = os.path.join(_HERE, "_build")

_f64p = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
_i64p = np.ctypeslib.ndpointer(dtype=np.int64, flags="C_CONTIGUOUS")
_i32p = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")


def build(force: bool = False) -> None:
    """Compile the oracle with the committed Makefile (gcc only)."""
    if force:
        subprocess.run(["make", "-C", _HERE, "clean"], check=True, capture_output=True)
    subprocess.run(["make", "-C", _HERE], check=True, capture_output=True)


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        path = os.path.join(_BUILD, "libnekoracle.so")
        src = os.path.join(_HERE, "nek_oracle.c")
        if not os.path.exists(path) or os.path.getmtime(path) < os.path.getmtime(src):
            build()
        L = C.CDLL(path)
        L.nko_zwgll.argtypes = [_f64p, _f64p, C.c_int]
        L.nko_dgll.argtypes = [_f64p, _f64p, _f64p, C.c_int]
        L.nko_mxm.argtypes = [_f64p, C.c_int, _f64p, C.c_int, _f64p, C.c_int]
        L.nko_box_mesh.argtypes = [C.c_int, C.c_int, C.c_int, _f64p, _f64p, _i32p, _f64p, _f64p, _f64p, _i64p]
        L.nko_xyzlin.argtypes = [C.c_int, C.c_int64, _f64p, _f64p, _f64p, _f64p, _f64p, _f64p, _f64p]
        L.nko_rescale_x.argtypes = [_f64p, C.c_int64, C.c_double, C.c_double]
        L.nko_geom_core.argtypes = [C.c_int, C.c_int64, _f64p, _f64p, _f64p, _f64p, _f64p, _f64p] + [_f64p] * 8
        L.nko_geodatstd.argtypes = [C.c_int, C.c_int64, _f64p, _f64p, _f64p, _f64p, _f64p, _f64p, _f64p]
        L.nko_setvert3d.argtypes = [_i64p, C.c_int, C.c_int64, _i64p, C.c_int]
        L.nko_setvert3d.restype = C.c_int64
        L.nko_check_p_bc.argtypes = [_i64p, C.c_int, C.c_int64, _i32p]
        L.nko_gs_op.argtypes = [_f64p, _i64p, C.c_int64, C.c_int]
        L.nko_rand_fld.argtypes = [_f64p, C.c_int64]
        L.nko_axhm1_bp5.argtypes = [_f64p, _f64p, _f64p, C.c_int, C.c_int64, _f64p, _f64p]
        L.nko_axhm1_bp5.restype = C.c_double
        L.nko_axhelm.argtypes = [_f64p, _f64p, _f64p, _f64p, C.c_int, C.c_int, C.c_int64, _f64p, _f64p] + [_f64p] * 7 + [C.c_void_p]
        L.nko_setprec_local.argtypes = [_f64p, _f64p, _f64p, C.c_int, C.c_int64, _f64p] + [_f64p] * 7 + [C.c_void_p]
        L.nko_cggos_bp5.argtypes = [_f64p, _f64p, _f64p, _f64p, _f64p, _i64p, _f64p, C.c_int, C.c_int64, _f64p, _f64p,
                                    C.c_double, C.c_int, C.c_void_p]
        L.nko_cggos_bp5.restype = C.c_int
        L.nko_cggo.argtypes = [_f64p, _f64p, _f64p, _f64p, _f64p, _f64p, _f64p, _i64p, C.c_int, C.c_int64, _f64p, _f64p] + \
            [_f64p] * 7 + [C.c_void_p, C.c_double, C.c_int, C.c_int, C.c_void_p]
        L.nko_cggo.restype = C.c_int
        L.nko_glrdif.argtypes = [_f64p, _f64p, C.c_int64]
        L.nko_glrdif.restype = C.c_double
        _lib = L
    return _lib


# ----------------------------------------------------------------------------- speclib
def zwgll(nx: int):
    """GLL points and weights (core/speclib.f:107 ZWGLL)."""
    z = np.zeros(nx)
    w = np.zeros(nx)
    lib().nko_zwgll(z, w, nx)
    return z, w


def dgll(z: np.ndarray):
    """Derivative matrix D and D^T, each as a Fortran-order (nx,nx) array: D[i,j] = D(i,j)
    (core/speclib.f:800 DGLL)."""
    nx = len(z)
    d = np.zeros(nx * nx)
    dt = np.zeros(nx * nx)
    lib().nko_dgll(d, dt, np.ascontiguousarray(z), nx)
    return d.reshape(nx, nx, order="F"), dt.reshape(nx, nx, order="F")


def _flat(a):
    """Fortran-order flattening to a contiguous float64 vector."""
    return np.ascontiguousarray(np.asarray(a, dtype=np.float64).ravel(order="F"))


def mxm(a: np.ndarray, b: np.ndarray):
    """C = A B with the reference summation order (core/mxm_std.f mxf*)."""
    n1, n2 = a.shape
    n3 = b.shape[1]
    c = np.zeros(n1 * n3)
    lib().nko_mxm(_flat(a), n1, _flat(b), n2, c, n3)
    return c.reshape(n1, n3, order="F")


# ----------------------------------------------------------------------------- case builder
class Case:
    """Everything the hot path consumes for one box mesh, computed by the oracle.

    Arrays are flat float64/int64 vectors in Nek's memory order u(i,j,k,e).
    """

    def __init__(self, nelx, nely, nelz, nx=8, periodic=(0, 0, 0), dirichlet=(1, 1, 1, 1, 1, 1), np_ranks=1,
                 lo=(0.0, 0.0, 0.0), hi=(1.0, 1.0, 1.0), deform=0.0, rescale=True, vertex_map=None):
        """rescale: True = usrdat2 of bp5.usr:40-44 (rescale_x to [0,1]), (a, b) = rescale_x to [a,b] (ethier.usr:224-228),
        False = none.  vertex_map(xc, yc, zc) -> (xc, yc, zc) moves the element vertices before the GLL points are
        generated, the way a case's usrdat does (examples/turbChannel/turbChannel.usr:334-358)."""
        L = lib()
        self.nx, self.nel = nx, nelx * nely * nelz
        self.nelx, self.nely, self.nelz = nelx, nely, nelz
        self.periodic, self.dirichlet = tuple(periodic), tuple(dirichlet)
        self.nxyz = nx ** 3
        self.n = self.nxyz * self.nel
        E = self.nel
        self.z, self.w = zwgll(nx)
        self.D, self.Dt = dgll(self.z)
        self.d, self.dt = _flat(self.D), _flat(self.Dt)
        self.w3 = _flat(np.einsum("i,j,k->ijk", self.w, self.w, self.w))  # coef.f:263-267
        per = np.asarray(periodic, dtype=np.int32)
        self.xc, self.yc, self.zc = np.zeros(8 * E), np.zeros(8 * E), np.zeros(8 * E)
        self.vertex = np.zeros(8 * E, dtype=np.int64)
        L.nko_box_mesh(nelx, nely, nelz, np.asarray(lo, dtype=np.float64), np.asarray(hi, dtype=np.float64), per,
                       self.xc, self.yc, self.zc, self.vertex)
        if vertex_map is not None:
            self.xc, self.yc, self.zc = (np.ascontiguousarray(a, dtype=np.float64) for a in vertex_map(self.xc, self.yc, self.zc))
        self.xm1, self.ym1, self.zm1 = np.zeros(self.n), np.zeros(self.n), np.zeros(self.n)
        L.nko_xyzlin(nx, E, self.z, self.xc, self.yc, self.zc, self.xm1, self.ym1, self.zm1)
        if rescale:  # bp5.usr:40-44 usrdat2
            ra, rb = (0.0, 1.0) if rescale is True else rescale
            for a in (self.xm1, self.ym1, self.zm1):
                L.nko_rescale_x(a, self.n, float(ra), float(rb))
        if deform:  # smooth, continuous deformation so all six factors are exercised
            x, y, z = self.xm1.copy(), self.ym1.copy(), self.zm1.copy()
            s = np.sin(np.pi * x) * np.sin(np.pi * y) * np.sin(np.pi * z)
            self.xm1 = x + deform * s
            self.ym1 = y + 0.7 * deform * s
            self.zm1 = z - 0.5 * deform * s
        # numbering
        self.glo_num = np.zeros(self.n, dtype=np.int64)
        self.ngv = L.nko_setvert3d(self.glo_num, nx, E, self.vertex, np_ranks)
        # mask (homogeneous Dirichlet on flagged box sides: bdry.f bcmask for 'v  '/'W  ')
        m = np.ones((nelz, nely, nelx, nx, nx, nx))  # [ez,ey,ex,k,j,i]
        dflag = list(dirichlet)
        if dflag[0] and not per[0]: m[:, :, 0, :, :, 0] = 0
        if dflag[1] and not per[0]: m[:, :, -1, :, :, -1] = 0
        if dflag[2] and not per[1]: m[:, 0, :, :, 0, :] = 0
        if dflag[3] and not per[1]: m[:, -1, :, :, -1, :] = 0
        if dflag[4] and not per[2]: m[0, :, :, 0, :, :] = 0
        if dflag[5] and not per[2]: m[-1, :, :, -1, :, :] = 0
        self.mask = np.ascontiguousarray(m.reshape(-1))
        # multiplicity (connect1.f:124-135 vmult = 1/dssum(1))
        self.mult = self.dssum(np.ones(self.n))
        self.mult = 1.0 / self.mult
        self._geom = None
        self._gf = None

    # -- geometry --------------------------------------------------------------------
    def geom(self):
        """g1..g6 (core order rr,ss,tt,rs,rt,st), bm1, jacm1 (coef.f:555-784)."""
        if self._geom is None:
            outs = [np.zeros(self.n) for _ in range(8)]
            lib().nko_geom_core(self.nx, self.nel, self.d, self.dt, self.w3, self.xm1, self.ym1, self.zm1, *outs)
            self._geom = outs
        return self._geom

    def gf(self):
        """BP5 interleaved factors gf(6,nxyz,E) (bp5.usr:623-699 geodatstd)."""
        if self._gf is None:
            g = np.zeros(6 * self.n)
            lib().nko_geodatstd(self.nx, self.nel, self.d, self.dt, self.w3, self.xm1, self.ym1, self.zm1, g)
            self._gf = g
        return self._gf

    def bm1(self):
        return self.geom()[6]

    def binv(self):
        """binvm1 = 1/dssum(bm1) (core/coef.f setinvm)."""
        return 1.0 / self.dssum(self.bm1().copy())

    # -- operators ---------------------------------------------------------------------
    def dssum(self, u, op=1):
        u = np.ascontiguousarray(u, dtype=np.float64).copy()
        lib().nko_gs_op(u, self.glo_num, self.n, op)
        return u

    def rand_fld_h1(self):
        """navier5.f:2687 rand_fld_h1 incl. dsavg (ic.f:1871)."""
        x = np.zeros(self.n)
        lib().nko_rand_fld(x, self.n)
        return self.dssum(x) * self.mult

    def ax_bp5(self, p):
        ap = np.zeros(self.n)
        pap = lib().nko_axhm1_bp5(ap, np.ascontiguousarray(p), self.gf(), self.nx, self.nel, self.d, self.dt)
        return ap, pap

    def axhelm(self, u, h1, h2, ifdfrm=None):
        g = self.geom()
        au = np.zeros(self.n)
        ifh2 = int(np.abs(h2).max() > 0)
        fp = None if ifdfrm is None else np.ascontiguousarray(ifdfrm, dtype=np.int32).ctypes.data
        lib().nko_axhelm(au, np.ascontiguousarray(u), np.ascontiguousarray(h1), np.ascontiguousarray(h2), ifh2,
                         self.nx, self.nel, self.d, self.dt, *g[:7], fp)
        return au

    def setprec(self, h1, h2, ifdfrm=None):
        """hmholtz.f:380-524 incl. dssum + invcol1."""
        g = self.geom()
        dp = np.zeros(self.n)
        fp = None if ifdfrm is None else np.ascontiguousarray(ifdfrm, dtype=np.int32).ctypes.data
        lib().nko_setprec_local(dp, np.ascontiguousarray(h1), np.ascontiguousarray(h2), self.nx, self.nel, self.dt,
                                *g[:7], fp)
        return 1.0 / self.dssum(dp)

    def cggos(self, rhs, x1, tol=-1e-8, maxit=500, history=False):
        u = np.zeros(self.n)
        hist = np.zeros(4 * maxit) if history else None
        it = lib().nko_cggos_bp5(u, np.ascontiguousarray(rhs), np.ascontiguousarray(x1), self.mult, self.mask,
                                 self.glo_num, self.gf(), self.nx, self.nel, self.d, self.dt, tol, maxit,
                                 None if hist is None else hist.ctypes.data)
        if history:
            return u, it, hist.reshape(maxit, 4)[:it]
        return u, it

    def cggo(self, f, h1, h2, mask=None, tin=1e-8, maxit=100, istep=1, ifdfrm=None, history=False):
        g = self.geom()
        mask = self.mask if mask is None else mask
        x = np.zeros(self.n)
        hist = np.zeros(3 * max(maxit, 1)) if history else None
        fp = None if ifdfrm is None else np.ascontiguousarray(ifdfrm, dtype=np.int32).ctypes.data
        it = lib().nko_cggo(x, np.ascontiguousarray(f), np.ascontiguousarray(h1), np.ascontiguousarray(h2),
                            np.ascontiguousarray(mask), self.mult, self.binv(), self.glo_num, self.nx, self.nel,
                            self.d, self.dt, *g[:7], fp, tin, maxit, istep,
                            None if hist is None else hist.ctypes.data)
        if history:
            return x, it, hist.reshape(-1, 3)[:max(it + 1, 1)]
        return x, it

    def bp5_problem(self):
        """e1, r1 of bp5.usr:352-360: e1 = mask*dsavg(rand), r1 = mask*dssum(A e1)."""
        e1 = self.rand_fld_h1() * self.mask
        ap, _ = self.ax_bp5(e1)
        r1 = self.dssum(ap) * self.mask
        return e1, r1


def glrdif(x, y):
    return lib().nko_glrdif(np.ascontiguousarray(x), np.ascontiguousarray(y), len(x))


# ----------------------------------------------------------------------------- CPU baseline
_cpu = None


def cpu_lib() -> C.CDLL:
    """OpenMP baseline, compiled -march=native ON THE RUNNING HOST (the prebuilt .so may have
    been built on a different CPU), falling back to the prebuilt one."""
    global _cpu
    if _cpu is None:
        src = os.path.join(_HERE, "bp5_cpu.c")
        path = None
        try:
            tmp = os.path.join(tempfile.gettempdir(), f"libbp5cpu_{os.getuid()}_{int(os.path.getmtime(src))}.so")
            if not os.path.exists(tmp):
                cc = "/usr/bin/gcc" if os.path.exists("/usr/bin/gcc") else "gcc"
                subprocess.run([cc, "-O3", "-march=native", "-fopenmp", "-fPIC", "-shared", "-std=gnu99",
                                "-o", tmp, src, "-lm"], check=True, capture_output=True)
            path = tmp
        except Exception:
            path = os.path.join(_BUILD, "libbp5cpu.so")
        L = C.CDLL(path)
        L.nkb_cpu_cggos.argtypes = [_f64p, _f64p, _f64p, _f64p, _i64p, _i32p, C.c_int64, _f64p, C.c_int, C.c_int64,
                                    _f64p, _f64p, C.c_int, C.c_int]
        L.nkb_cpu_cggos.restype = C.c_double
        L.nkb_cpu_max_threads.restype = C.c_int
        _cpu = L
    return _cpu


def gs_groups(glo_num: np.ndarray):
    """CSR of id groups (size >= 2) for the shared-memory gs of the CPU baseline."""
    nz = np.nonzero(glo_num)[0]
    order = nz[np.argsort(glo_num[nz], kind="stable")]
    ids = glo_num[order]
    starts = np.flatnonzero(np.r_[True, ids[1:] != ids[:-1]])
    sizes = np.diff(np.r_[starts, len(ids)])
    keep = sizes >= 2
    sel = np.repeat(keep, sizes)
    idx = order[sel].astype(np.int32)
    off = np.r_[0, np.cumsum(sizes[keep])].astype(np.int64)
    return off, idx


def cpu_cggos(case: "Case", rhs, maxit: int, nthreads: int = 0):
    """Time `maxit` fixed iterations of the restated cggos on the host cores.
    Returns (u, seconds, threads)."""
    L = cpu_lib()
    off, idx = gs_groups(case.glo_num)
    u = np.zeros(case.n)
    nt = nthreads or L.nkb_cpu_max_threads()
    sec = L.nkb_cpu_cggos(u, np.ascontiguousarray(rhs), case.mult, case.mask, off, idx, len(off) - 1, case.gf(),
                          case.nx, case.nel, case.d, case.dt, maxit, nt)
    return u, sec, nt


def bp5_partitioned_reference(nelx, nely, nelz, layout, rank, maxit=40, deform=0.04, nx=8):
    """The undivided BP5 solve a brick-partitioned run must reproduce on rank `rank` (checker for the multi-GPU tests and
    for bench.py's `parity` field at N > 1).  Every rank draws the seed-1 ran1 stream over its LOCAL nodes in local element
    order (core/navier5.f:2665-2698), so the exact solution of the undivided mesh is that stream placed brick by
    brick; e1 = dsavg(.)*mask, r1 = mask*dssum(A e1) (bp5.usr:352-360), then cggos on the whole mesh.
    Returns dict(take=<global node indices of this rank's nodes in its local order>, e1, r1, u, hist, it, mult, glo_num)
    with the fields already restricted to the rank."""
    px, py, pz = layout
    case = Case(nelx, nely, nelz, nx=nx, deform=deform)
    nxyz = nx ** 3
    lx, ly, lz = nelx // px, nely // py, nelz // pz
    eg = np.arange(case.nel)
    ex, ey, ez = eg % nelx, (eg // nelx) % nely, eg // (nelx * nely)
    owner = (ex // lx) + px * ((ey // ly) + py * (ez // lz))
    local_of = (ex % lx) + lx * ((ey % ly) + ly * (ez % lz))
    nloc = lx * ly * lz * nxyz
    stream = np.zeros(nloc)
    lib().nko_rand_fld(stream, nloc)
    rnd = stream.reshape(-1, nxyz)[local_of].reshape(-1)
    e1 = case.dssum(rnd) * case.mult * case.mask
    ap, _ = case.ax_bp5(e1)
    r1 = case.dssum(ap) * case.mask
    uref, itref, href = case.cggos(r1, e1, maxit=maxit, history=True)
    mine = np.flatnonzero(owner == rank)
    order = mine[np.argsort(local_of[mine])]
    take = (order[:, None] * nxyz + np.arange(nxyz)[None, :]).reshape(-1)
    return dict(case=case, take=take, e1=e1[take], r1=r1[take], u=uref[take], hist=href, it=itref, mult=case.mult[take],
                glo_num=case.glo_num[take])
