"""ctypes front end of the transpiled reference (oracle/_ref/libnekref_*.so, built by oracle/ref_build.py).

TEST INFRASTRUCTURE ONLY -- see oracle/f77c.py.  `Ref` gives numpy views onto the reference's COMMON blocks by variable
name (Fortran order, the reference's own extents from the generated SIZE) and calls any translated routine with Fortran
calling conventions (everything by reference, hidden CHARACTER lengths appended).
"""
from __future__ import annotations

import ctypes as C
import json
import os

import numpy as np

from . import ref_build

_DT = {"r8": np.float64, "r4": np.float32, "i4": np.int32, "i8": np.int64, "l4": np.int32, "i2": np.int16, "i1": np.int8,
       "l1": np.int8}


def available(lx1=8, lx2=None, lelt=64) -> bool:
    try:
        ref_build.build(lx1, lx2, lelt)
        return True
    except Exception:
        return False


class Ref:
    _cache = {}

    def __new__(cls, lx1=8, lx2=None, lelt=64, lgmres=30, fresh=False, hybrid=False):
        """fresh=True loads a private copy of the library: the reference keeps state in SAVE variables (icalld counters,
        hmh_gmres' norm_fac, ...) and in the stand-ins' handle tables, which must not leak from one test case to the next."""
        key = (lx1, lx2 or lx1, lelt, lgmres)
        if hybrid:        # the drop-in variant (ref_build.HYB_STOP routines come from libnekb200.so): always a private copy
            self = super().__new__(cls)
            self._init(*key, fresh=True, hybrid=True)
            return self
        if fresh:
            self = super().__new__(cls)
            self._init(*key, fresh=True)
            return self
        if key not in cls._cache:
            self = super().__new__(cls)
            self._init(*key)
            cls._cache[key] = self
        return cls._cache[key]

    def _init(self, lx1, lx2, lelt, lgmres, fresh=False, hybrid=False):
        so = ref_build.build(lx1, lx2, lelt, lgmres, hybrid=hybrid)
        self.hybrid = hybrid
        if hybrid:
            # libnekb200.so first, globally (it carries the soname the hybrid's DT_NEEDED entry asks for, so the private copy
            # made below binds to this very instance)
            from nek5000_b200 import lib as _product
            _product()
        self.meta = json.load(open(so.replace("libnekref_", "nekref_").replace(".so", ".json")))
        if fresh:
            import shutil
            import tempfile
            fd, tmp = tempfile.mkstemp(suffix=".so", prefix="nekref_")
            os.close(fd)
            shutil.copyfile(so, tmp)
            self.lib = C.CDLL(tmp)      # a distinct path gives distinct globals
            os.unlink(tmp)
        else:
            self.lib = C.CDLL(so)
        self.lx1, self.lx2, self.lelt = lx1, lx2, lelt
        self._blocks = {}
        self._keep = []

    # ---- COMMON access --------------------------------------------------------------------------------------------
    def _block(self, blk):
        if blk not in self._blocks:
            size = self.meta["common_size"][blk]
            self._blocks[blk] = (C.c_char * size).in_dll(self.lib, "cb_" + blk)
        return self._blocks[blk]

    def var(self, name, unit=None) -> np.ndarray:
        """Writable numpy view of a COMMON variable (scalars: shape (1,); CHARACTER: numpy bytes array).  `unit` selects
        the view a particular routine declares when the name is placed differently elsewhere."""
        m = self.meta["commons_alt"].get(unit, {}).get(name.lower()) if unit else None
        m = m or self.meta["commons"][name.lower()]
        buf = self._block(m["block"])
        dims = m["dims"] or [1]
        cnt = int(np.prod(dims))
        if m["type"] == "ch":
            a = np.frombuffer(buf, dtype="S%d" % m["elsize"], count=cnt, offset=m["offset"])
        else:
            a = np.frombuffer(buf, dtype=_DT[m["type"]], count=cnt, offset=m["offset"])
        return a.reshape(dims, order="F")

    def lows(self, name):
        return self.meta["commons"][name.lower()]["lows"]

    def set(self, name, value):
        v = self.var(name)
        v[...] = value

    def get(self, name):
        v = self.var(name)
        return v[0] if v.shape == (1,) and not self.meta["commons"][name.lower()]["dims"] else v

    # ---- what the reference logs (write(6,...) in the routines of ref_build.TRACE_UNITS) -----------------------------------
    def trace(self, on=True):
        """Starts (and clears) or stops the recording of the numeric items of the traced routines' write(6,...) statements."""
        self.lib.nekref_trace_enable(1 if on else 0)

    def trace_records(self, unit=None, nvals=None):
        """[(unit, values)] recorded since trace(True); optionally only those of routine `unit` with `nvals` numeric items."""
        out = []
        name, vals = C.create_string_buffer(24), (C.c_double * 12)()
        for i in range(self.lib.nekref_trace_count()):
            n = self.lib.nekref_trace_get(i, name, vals)
            u = name.value.decode()
            if (unit is None or u == unit) and (nvals is None or n == nvals):
                out.append((u, np.array(vals[:n])))
        return out

    # ---- calls ----------------------------------------------------------------------------------------------------------
    def call(self, name, *args, restype=None):
        """Calls `name_` Fortran-style.  numpy arrays pass their buffer, ints/floats/bools pass a temporary by reference
        (returned values are visible through `Ref.out`), str passes characters + hidden length."""
        f = getattr(self.lib, name.lower() + "_")
        f.restype = {None: None, "r8": C.c_double, "i4": C.c_int, "i8": C.c_longlong, "l4": C.c_int}[restype]
        cargs, hidden, self.out = [], [], []
        for a in args:
            if isinstance(a, np.ndarray):
                assert a.flags.f_contiguous or a.flags.c_contiguous
                cargs.append(a.ctypes.data_as(C.c_void_p))
            elif isinstance(a, (bool, np.bool_)):
                t = C.c_int(1 if a else 0)
                self.out.append(t)
                cargs.append(C.byref(t))
            elif isinstance(a, (int, np.integer)):
                t = C.c_int(int(a))
                self.out.append(t)
                cargs.append(C.byref(t))
            elif isinstance(a, (float, np.floating)):
                t = C.c_double(float(a))
                self.out.append(t)
                cargs.append(C.byref(t))
            elif isinstance(a, str):
                b = a.encode()
                cargs.append(C.c_char_p(b))
                hidden.append(C.c_long(len(b)))
            elif isinstance(a, C._SimpleCData):
                self.out.append(a)
                cargs.append(C.byref(a))
            else:
                raise TypeError(f"argument {a!r}")
        return f(*cargs, *hidden)


class RefCase:
    """Initialises the transpiled reference for an oracle.Case box mesh by running the reference's own set-up routines
    in the order of nek_init (core/drive1.f:34-219): initdim, initdat, [mesh + BCs from the Case instead of readat],
    initds/dsset/setedge + setupds + multiplicity (connect1.f:43-135), genwz, geom1/geom2/volume/setinvm/setdef
    (gengeom, core/coef.f / drive2.f), bcmask."""

    def __init__(self, case, lelt=None, lx2=None, ifsplit=True, nfield=1, fresh=True, lgmres=30, hybrid=False, device=0):
        """hybrid=True: the drop-in variant -- the same set-up sequence, with the glue calls of INTEGRATION.md
        (oracle/hyb_glue.c) where a Nek5000 build would make them: nekb_init before the first gs_setup, the field handle
        after setupds, the registration of COMMON state after the geometry is complete."""
        nx, E = case.nx, case.nel
        lelt = lelt or max(64, E)
        R = self.R = Ref(nx, lx2 or nx, lelt, lgmres, fresh=fresh, hybrid=hybrid)
        self.case, self.E, self.nx = case, E, nx
        R.call("initdim")
        R.call("initdat")
        for k, v in dict(nelv=E, nelt=E, nelgv=E, nelgt=E, nelg=E, nfield=nfield, nid=0, np=1, mid=0, mp=1, nio=-1, if3d=1, ifaxis=0,
                         ifflow=1, ifheat=0, iftran=1, ifsplit=int(ifsplit), ifield=1, istep=0, nekreal=3, ifstrs=0,
                         ifmvbd=0, ifmhd=0, ifcvode=0, ifgmsh3=0, nelx=case.nelx, nely=case.nely, nelz=case.nelz).items():
            if k in R.meta["commons"]:
                R.set(k, v)
        R.var("nelfld")[:] = E
        R.var("param")[58] = 1.0               # param(59)=1: every element takes the general (deformed) branch
        R.var("lglel")[:E] = np.arange(1, E + 1)
        R.var("gllel")[:E] = np.arange(1, E + 1)
        R.var("gllnid")[:E] = 0
        # boundary conditions in preprocessor face order (1:y-,2:x+,3:y+,4:x-,5:z-,6:z+)
        cbc = R.var("cbc")
        cbc[...] = b"E  "
        per = getattr(case, "periodic", (0, 0, 0))
        dflag = getattr(case, "dirichlet", (1, 1, 1, 1, 1, 1))
        eid = np.arange(E)
        ex, ey, ez = eid % case.nelx, (eid // case.nelx) % case.nely, eid // (case.nelx * case.nely)
        for f, on, ax, di in ((4, ex == 0, 0, 0), (2, ex == case.nelx - 1, 0, 1), (1, ey == 0, 1, 2),
                              (3, ey == case.nely - 1, 1, 3), (5, ez == 0, 2, 4), (6, ez == case.nelz - 1, 2, 5)):
            cbc[f - 1, eid[on], 1] = b"P  " if per[ax] else {0: b"O  ", 1: b"v  ", 2: b"SYM"}[int(dflag[di])]
        # setlog (core/bdry.f:24,50-57): no outflow face anywhere -> the pressure has the constant null space
        R.set("ifvcor", int(not (cbc[:, :E, 1] == b"O  ").any()))
        sh = (nx, nx, nx, E)
        R.var("xc")[:, :E] = case.xc.reshape(8, E, order="F")
        R.var("yc")[:, :E] = case.yc.reshape(8, E, order="F")
        R.var("zc")[:, :E] = case.zc.reshape(8, E, order="F")
        R.var("vertex")[:8 * E] = case.vertex
        R.call("initds")
        R.call("dsset", 3, 3, 3)
        R.call("setedge")
        R.call("genwz")
        for n, a in (("xm1", case.xm1), ("ym1", case.ym1), ("zm1", case.zm1)):
            R.var(n)[..., :E] = a.reshape(sh, order="F")
        # numbering + gs handle + multiplicity (connect1.f:81-135)
        if hybrid:
            R.call("nekhyb_init", device)
        R.call("setupds", R.var("gsh_fld")[1:2], nx, nx, nx, E, E, R.var("vertex"), R.var("glo_num"))
        R.var("gsh_fld")[2] = R.var("gsh_fld")[1]
        if hybrid:
            R.call("nekhyb_set_field")
        n = nx ** 3 * E
        R.call("rone", R.var("vmult"), n)
        R.call("dssum", R.var("vmult"), nx, nx, nx)
        R.call("invcol1", R.var("vmult"), n)
        # geometry (drive2.f gengeom)
        R.call("geom1", R.var("xm1"), R.var("ym1"), R.var("zm1"))
        R.call("geom2")
        R.call("volume")
        R.call("setinvm")
        R.call("setdef")
        R.call("sfastax")
        R.call("bcmask")
        R.set("ifield", 1)
        if hybrid:
            R.call("nekhyb_register")

    def fld(self, name):
        """Flat copy (Nek memory order) of the first E elements of a field in COMMON."""
        v = self.R.var(name)
        if v.ndim == 1:
            return v[:v.size // self.R.lelt * self.E].copy()
        return v[..., :self.E].ravel(order="F").copy()
