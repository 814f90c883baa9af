"""Host-side mirror of the Nek5000 interfaces of the hot path, over the C-ABI of libnekb200.so.

The functions carry the reference's names, argument order and in-place semantics (numpy arrays stand in
for the Fortran arrays; everything is passed by reference exactly as gfortran would), so that a test
reads like a call site in the reference:

    axhelm(au,u,helm1,helm2,imesh,isd)                      core/hmholtz.f:72
    cggo(x,f,h1,h2,mask,mult,imsh,tin,maxit,isd,binv,name)  core/hmholtz.f:611
    setprec(dpcm1,helm1,helm2,imsh,isd)                     core/hmholtz.f:380
    dssum(u,nx,ny,nz) / dsop(u,op,nx,ny,nz)                 core/dssum.f:33,100
    setupds(gs_handle,nx,ny,nz,nel,melg,vertex,glo_num)     core/dssum.f:1
    fgslib_gs_setup/op/op_many/op_fields/free               gslib v1.0.9 (call sites core/dssum.f:20,79,198,277)
    cggos(u1,rhs1,x1,rmult,binv,tin,maxit,bpname)           examples/bp5/bp5.usr:797
    axhm1(pap,ap1,p1,h1,h2,bpname)                          examples/bp5/bp5.usr:1389
    glsc3(a,b,mult,n)                                       core/math.f:775
    h1mg_setup() / h1mg_solve(z,rhs,if_hybrid)              core/hsmg.f:2234,1855
    hmh_gmres(res,h1,h2,wt,iter)                            core/gmres.f:304

State that the Fortran routines read from COMMON blocks is registered with the set_* functions
(include/nekb200.h section B).  All compute runs in hand-written CUDA kernels; a missing library or GPU is
an error, never a fallback.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from ._lib import ALLGATHER_FN, ALLTOALLV_FN, NekbError, check, lib

__all__ = [
    "NekbError", "init", "finalize", "launch_count", "set_nel", "set_gll", "set_dxyz", "set_geom", "set_geom_bp5",
    "set_geom_from_xyz", "get_geom", "set_ifdfrm", "set_v1mask", "set_ifield", "set_field_handle", "set_step_info",
    "niterhm", "setvert3d", "setupds", "fgslib_gs_setup", "fgslib_gs_op", "fgslib_gs_op_many", "fgslib_gs_op_fields",
    "fgslib_gs_free", "gs_get_map", "gs_info", "dssum", "dsop", "axhelm", "setprec", "cggo", "cggos", "axhm1", "glsc3",
    "DevArray", "set_transport_torch", "comm_init_torch", "h1mg_setup", "h1mg_solve", "h1mg_info", "h1mg_get", "h1mg_free",
    "hsmg_setup", "hsmg_solve", "local_solves_fdm", "hsmg_get", "set_pressure_state", "hmh_gmres", "hmholtz", "set_param", "set_binv", "fdm_h1_setup", "set_kfldfdm", "set_fdm_prec_h1b", "fdm_h1", "fdm_h1_get",
]

_state = {"lx1": 0, "nelt": 0, "np": 1, "keep": []}


def _i(v):
    return C.byref(C.c_int(int(v)))


def _ptr(a: np.ndarray):
    assert a.flags["C_CONTIGUOUS"] and a.dtype in (np.float64, np.int64, np.int32), "contiguous f64/i64/i32 arrays only"
    return C.c_void_p(a.ctypes.data)


# --------------------------------------------------------------------------------------------- lifecycle
def init(device: int = 0, lx1: int = 8, ldim: int = 3) -> None:
    check(lib().nekb_init(device, lx1, ldim))
    _state["lx1"] = lx1


def finalize() -> None:
    lib().nekb_finalize()
    _state.update(lx1=0, nelt=0, np=1)


def launch_count(reset: bool = False) -> int:
    return int(lib().nekb_launch_count(1 if reset else 0))


# --------------------------------------------------------------------------------------------- registration
def set_nel(nelv: int, nelt: int | None = None) -> None:
    nelt = nelv if nelt is None else nelt
    check(lib().nekb_set_nel(nelv, nelt))
    _state["nelt"] = nelt


def set_gll(zgm1, wxm1) -> None:
    check(lib().nekb_set_gll(np.ascontiguousarray(zgm1, dtype=np.float64), np.ascontiguousarray(wxm1, dtype=np.float64)))


def set_dxyz(dxm1, dxtm1=None) -> None:
    """dxm1(lx1,lx1) as a 2-D array indexed [i,j] = dxm1(i,j) (it is flattened in Fortran order)."""
    d = np.ascontiguousarray(np.asarray(dxm1, dtype=np.float64).ravel(order="F"))
    dt = np.ascontiguousarray((np.asarray(dxm1).T if dxtm1 is None else np.asarray(dxtm1, dtype=np.float64)).ravel(order="F"))
    check(lib().nekb_set_dxyz(d, dt))


def set_geom(g1m1, g2m1, g3m1, g4m1, g5m1, g6m1, bm1) -> None:
    check(lib().nekb_set_geom(*[np.ascontiguousarray(a, dtype=np.float64) for a in (g1m1, g2m1, g3m1, g4m1, g5m1, g6m1, bm1)]))


def set_geom_bp5(gf) -> None:
    check(lib().nekb_set_geom_bp5(np.ascontiguousarray(gf, dtype=np.float64)))


def set_geom_from_xyz(xm1, ym1, zm1, bp5_form: bool = False) -> None:
    check(lib().nekb_set_geom_from_xyz(*[np.ascontiguousarray(a, dtype=np.float64) for a in (xm1, ym1, zm1)], int(bp5_form)))


def get_geom(gf: bool = False):
    """Returns (g1m1..g6m1, bm1) or, with gf=True, the interleaved gf(6,nxyz,nelt) as a flat vector."""
    n = _state["lx1"] ** 3 * _state["nelt"]
    if gf:
        out = np.zeros(6 * n)
        check(lib().nekb_get_geom(None, None, None, None, None, None, None, _ptr(out)))
        return out
    outs = [np.zeros(n) for _ in range(7)]
    check(lib().nekb_get_geom(*[_ptr(a) for a in outs], None))
    return outs


def set_ifdfrm(ifdfrm) -> None:
    if ifdfrm is None:
        check(lib().nekb_set_ifdfrm(None))
    else:
        a = np.ascontiguousarray(ifdfrm, dtype=np.int32)
        check(lib().nekb_set_ifdfrm(_ptr(a)))


def set_v1mask(v1mask) -> None:
    check(lib().nekb_set_v1mask(np.ascontiguousarray(v1mask, dtype=np.float64)))


def set_ifield(ifield: int) -> None:
    check(lib().nekb_set_ifield(ifield))


def set_field_handle(ifield: int, gs_handle: int) -> None:
    check(lib().nekb_set_field_handle(ifield, gs_handle))


def last_history() -> np.ndarray:
    """History of the most recent cggo solve, shape (checks, 3): rtz1, rbn2, rho; or of the most recent hmh_gmres /
    hmh_flex_cg solve, shape (iterations, 1): rnorm."""
    rows, cols = C.c_int(0), C.c_int(0)
    check(lib().nekb_last_history(None, 0, C.byref(rows), C.byref(cols)))
    out = np.zeros((rows.value, max(cols.value, 1)))
    if out.size:
        check(lib().nekb_last_history(out.ctypes.data, out.size, C.byref(rows), C.byref(cols)))
    return out


def set_restol(ifield: int, restol: float) -> None:
    """TSTEP restol(ifield): overrules cggo's tolerance when non-zero (core/hmholtz.f:676)."""
    check(lib().nekb_set_restol(ifield, restol))


def set_step_info(istep: int, volvm1: float, voltm1: float | None = None) -> None:
    check(lib().nekb_set_step_info(istep, volvm1, volvm1 if voltm1 is None else voltm1))


def niterhm() -> int:
    return int(lib().nekb_niterhm())


# --------------------------------------------------------------------------------------------- mesh files (host only)
def re2_info(path: str) -> dict:
    """core/reader_re2.f:543-639 header + section table of a .re2 file."""
    nelgt, nelgv, ncurve = C.c_int64(0), C.c_int64(0), C.c_int64(0)
    ldim, wd, nsec = C.c_int(0), C.c_int(0), C.c_int(0)
    nbc = np.zeros(16, dtype=np.int64)
    check(lib().nekb_re2_info(path.encode(), C.byref(nelgt), C.byref(ldim), C.byref(nelgv), C.byref(wd), C.byref(ncurve),
                              C.byref(nsec), _ptr(nbc), 16))
    return dict(nelgt=int(nelgt.value), ldim=int(ldim.value), nelgv=int(nelgv.value), wdsize=int(wd.value),
                ncurve=int(ncurve.value), nbc=[int(v) for v in nbc[:nsec.value]])


def re2_read_mesh(path: str, e0: int = 0, nel: int | None = None):
    """Elements [e0, e0+nel) -> xc, yc, zc (nel, 2^ldim; preprocessor corner order), igroup."""
    info = re2_info(path)
    nel = info["nelgt"] - e0 if nel is None else nel
    nv = 1 << info["ldim"]
    xc, yc, zc = (np.zeros((nel, nv)) for _ in range(3))
    grp = np.zeros(nel, dtype=np.int32)
    check(lib().nekb_re2_read_mesh(path.encode(), e0, nel, _ptr(xc), _ptr(yc), _ptr(zc), _ptr(grp)))
    return xc, yc, zc, grp


def re2_read_bc(path: str, section: int = 0):
    """One boundary-condition section -> cbc (nelgt, 6) of 3-character codes (blank where the file has no record),
    bc (nelgt, 6, 5)."""
    info = re2_info(path)
    cbc = np.full((info["nelgt"], 6), b"   ", dtype="S3")
    bc = np.zeros((info["nelgt"], 6, 5))
    check(lib().nekb_re2_read_bc(path.encode(), section, C.c_void_p(cbc.ctypes.data), _ptr(bc)))
    return cbc, bc


def re2_read_curves(path: str):
    """The curved-side section -> ccurve (nelgt, 12) of 1-character codes (blank where the file has no record),
    curve (nelgt, 12, 5)  (core/reader_re2.f:160-290, buf_to_curve)."""
    info = re2_info(path)
    ccurve = np.full((info["nelgt"], 12), b" ", dtype="S1")
    curve = np.zeros((info["nelgt"], 12, 5))
    check(lib().nekb_re2_read_curves(path.encode(), C.c_void_p(ccurve.ctypes.data), _ptr(curve)))
    return ccurve, curve


def ma2_read(path: str, nlv: int = 8, e0: int = 0, nel: int | None = None):
    """core/map2.f:712-941: (header[7], leaf[nel], vertex[nel, nlv]) of a .ma2 file."""
    n = C.c_int64(0)
    hdr = np.zeros(7, dtype=np.int64)
    check(lib().nekb_ma2_info(path.encode(), C.byref(n), _ptr(hdr)))
    nel = int(n.value) - e0 if nel is None else nel
    leaf = np.zeros(nel, dtype=np.int32)
    vertex = np.zeros((nel, nlv), dtype=np.int64)
    check(lib().nekb_ma2_read(path.encode(), nlv, e0, nel, _ptr(leaf), _ptr(vertex)))
    return hdr, leaf, vertex


def co2_read(path: str, nlv: int = 8, e0: int = 0, nel: int | None = None):
    """core/map2.f:338-473 read_con: (nelgt, nelgv, eid[nel], vertex[nel, nlv]) of a .co2 connectivity file."""
    ngt, ngv, nv = C.c_int64(0), C.c_int64(0), C.c_int(0)
    check(lib().nekb_co2_info(path.encode(), C.byref(ngt), C.byref(ngv), C.byref(nv)))
    nel = int(ngt.value) - e0 if nel is None else nel
    eid = np.zeros(nel, dtype=np.int64)
    vertex = np.zeros((nel, nlv), dtype=np.int64)
    check(lib().nekb_co2_read(path.encode(), nlv, e0, nel, _ptr(eid), _ptr(vertex)))
    return int(ngt.value), int(ngv.value), eid, vertex


def assign_gllnid(leaf, nelgv: int | None = None, np_ranks: int = 1) -> np.ndarray:
    """core/map2.f:943-1026 assign_gllnid: RSB leaves -> 0-based rank of every global element."""
    g = np.ascontiguousarray(leaf, dtype=np.int32).copy()
    check(lib().nekb_assign_gllnid(_ptr(g), len(g), len(g) if nelgv is None else nelgv, np_ranks))
    return g


# --------------------------------------------------------------------------------------------- numbering / gs
def setvert3d(nx: int, nel: int, vertex, np_ranks: int = 1):
    """core/navier8.f:2004 setvert3d -> (glo_num, ngv).  Host only (no GPU needed)."""
    v = np.ascontiguousarray(vertex, dtype=np.int64).reshape(-1)
    glo = np.zeros(nx ** 3 * nel, dtype=np.int64)
    ngv = C.c_int64(0)
    check(lib().nekb_setvert3d(glo, C.byref(ngv), nx, nel, v, np_ranks))
    return glo, int(ngv.value)


def setupds(nx: int, nel: int, vertex, melg: int | None = None):
    """core/dssum.f:1 setupds -> (gs_handle, glo_num)."""
    v = np.ascontiguousarray(vertex, dtype=np.int64).reshape(-1)
    glo = np.zeros(nx ** 3 * nel, dtype=np.int64)
    h = C.c_int(-1)
    lib().setupds_(C.byref(h), _i(nx), _i(nx), _i(nx), _i(nel), _i(melg or nel), v, glo)
    return int(h.value), glo


def fgslib_gs_setup(ids, comm: int = 0, np_ranks: int | None = None) -> int:
    ids = np.ascontiguousarray(ids, dtype=np.int64)
    h = C.c_int(-1)
    lib().fgslib_gs_setup_(C.byref(h), ids, _i(len(ids)), _i(comm), _i(_state["np"] if np_ranks is None else np_ranks))
    return int(h.value)


def fgslib_gs_op(handle: int, u: np.ndarray, dom: int = 1, op: int = 1, transpose: int = 0) -> None:
    lib().fgslib_gs_op_(_i(handle), _ptr(u), _i(dom), _i(op), _i(transpose))


def fgslib_gs_op_many(handle: int, us, dom: int = 1, op: int = 1, transpose: int = 0) -> None:
    ptrs = [_ptr(u) for u in us] + [None] * (6 - len(us))
    lib().fgslib_gs_op_many_(_i(handle), *ptrs, _i(len(us)), _i(dom), _i(op), _i(transpose))


def fgslib_gs_op_fields(handle: int, u: np.ndarray, stride: int, n: int, dom: int = 1, op: int = 1, transpose: int = 0) -> None:
    lib().fgslib_gs_op_fields_(_i(handle), _ptr(u), _i(stride), _i(n), _i(dom), _i(op), _i(transpose))


def fgslib_gs_free(handle: int) -> None:
    lib().fgslib_gs_free_(_i(handle))


def gs_info(handle: int):
    a, b, c = C.c_int64(0), C.c_int64(0), C.c_int64(0)
    check(lib().nekb_gs_info(handle, C.byref(a), C.byref(b), C.byref(c)))
    return int(a.value), int(b.value), int(c.value)


def gs_get_map(handle: int):
    """CSR (offsets int64, member indices int32) of the handle's local gather-scatter map."""
    ng, nm, _ = gs_info(handle)
    off, idx = np.zeros(ng + 1, dtype=np.int64), np.zeros(max(nm, 1), dtype=np.int32)
    check(lib().nekb_gs_get_map(handle, off, idx))
    return off, idx[:nm]


def gs_get_remote(handle: int):
    npeers, nitems = C.c_int(0), C.c_int64(0)
    check(lib().nekb_gs_remote_info(handle, C.byref(npeers), C.byref(nitems)))
    peers = np.zeros(max(npeers.value, 1), dtype=np.int32)
    off = np.zeros(npeers.value + 1, dtype=np.int64)
    rep = np.zeros(max(nitems.value, 1), dtype=np.int32)
    if npeers.value:
        check(lib().nekb_gs_get_remote(handle, peers, off, rep))
    return peers[:npeers.value], off, rep[:nitems.value]


def gs_discover(uniq_ids):
    """Host-only rendezvous of shared ids over the registered transport -> (peers, peer_off, item_ids)."""
    u = np.ascontiguousarray(uniq_ids, dtype=np.int64)
    npeers, nitems = C.c_int(0), C.c_int64(0)
    check(lib().nekb_gs_discover(u, len(u), C.byref(npeers), C.byref(nitems), None, None, None))
    peers = np.zeros(max(npeers.value, 1), dtype=np.int32)
    off = np.zeros(npeers.value + 1, dtype=np.int64)
    ids = np.zeros(max(nitems.value, 1), dtype=np.int64)
    check(lib().nekb_gs_discover(u, len(u), C.byref(npeers), C.byref(nitems), _ptr(peers), _ptr(off), _ptr(ids)))
    return peers[:npeers.value], off, ids[:nitems.value]


def dssum(u: np.ndarray, nx: int | None = None, ny: int | None = None, nz: int | None = None) -> None:
    nx = nx or _state["lx1"]
    lib().dssum_(_ptr(u), _i(nx), _i(ny or nx), _i(nz or nx))


def dsop(u: np.ndarray, op: str, nx: int | None = None, ny: int | None = None, nz: int | None = None) -> None:
    nx = nx or _state["lx1"]
    o = op.ljust(3).encode()
    lib().dsop_(_ptr(u), o, _i(nx), _i(ny or nx), _i(nz or nx), 3)


# --------------------------------------------------------------------------------------------- operators / solvers
def vec_dssum(u, v, w, nx: int | None = None, ny: int | None = None, nz: int | None = None) -> None:
    """core/dssum.f:163 vec_dssum."""
    n = _state["lx1"]
    lib().vec_dssum_(_ptr(u), _ptr(v), _ptr(w), _i(nx or n), _i(ny or n), _i(nz or n))


def vec_dsop(u, v, w, op: str, nx: int | None = None, ny: int | None = None, nz: int | None = None) -> None:
    """core/dssum.f:198 vec_dsop."""
    n = _state["lx1"]
    o = op.encode().ljust(3)[:3]
    lib().vec_dsop_(_ptr(u), _ptr(v), _ptr(w), _i(nx or n), _i(ny or n), _i(nz or n), o, len(o))


def nvec_dssum(u, stride: int, n: int, gs_handle: int) -> None:
    """core/dssum.f:260 nvec_dssum."""
    lib().nvec_dssum_(_ptr(u), _i(stride), _i(n), _i(gs_handle))


def dsavg(u) -> None:
    """core/ic.f:1871 dsavg."""
    lib().dsavg_(_ptr(u))


def axhelm(au: np.ndarray, u: np.ndarray, helm1: np.ndarray, helm2: np.ndarray, imesh: int = 1, isd: int = 1) -> None:
    lib().axhelm_(_ptr(au), _ptr(u), _ptr(helm1), _ptr(helm2), _i(imesh), _i(isd))


def setprec(dpcm1: np.ndarray, helm1: np.ndarray, helm2: np.ndarray, imsh: int = 1, isd: int = 1) -> None:
    lib().setprec_(_ptr(dpcm1), _ptr(helm1), _ptr(helm2), _i(imsh), _i(isd))


def cggo(x, f, h1, h2, mask, mult, imsh, tin, maxit, isd, binv, name: str = "VELX") -> int:
    nm = name.ljust(4).encode()
    lib().cggo_(_ptr(x), _ptr(f), _ptr(h1), _ptr(h2), _ptr(mask), _ptr(mult), _i(imsh), C.byref(C.c_double(tin)), _i(maxit),
                _i(isd), _ptr(binv), nm, 4)
    return niterhm()


def hmholtz(name: str, u, rhs, h1, h2, mask, mult, imsh: int, tli: float, maxit: int, isd: int = 1) -> int:
    """core/hmholtz.f:2 hmholtz(name,u,rhs,h1,h2,mask,mult,imsh,tli,maxit,isd); rhs is dssum'ed and masked in place."""
    nm = name.ljust(4).encode()
    lib().hmholtz_(nm, _ptr(u), _ptr(rhs), _ptr(h1), _ptr(h2), _ptr(mask), _ptr(mult), _i(imsh), C.byref(C.c_double(tli)),
                   _i(maxit), _i(isd), 4)
    return niterhm()


def set_velocity_state(v1mask, v2mask, v3mask, vmult) -> None:
    """COMMON state ophinv reads: core/SOLN v1mask, v2mask, v3mask, vmult."""
    a = [np.ascontiguousarray(q, dtype=np.float64).reshape(-1) for q in (v1mask, v2mask, v3mask, vmult)]
    check(lib().nekb_set_velocity_state(*[_ptr(q) for q in a]))


def ophinv(o1, o2, o3, i1, i2, i3, h1, h2, tolh: float, nmxhi: int):
    """core/induct.f:1022 ophinv(o1,o2,o3,i1,i2,i3,h1,h2,tolh,nmxhi); returns the three iteration counts."""
    lib().ophinv_(_ptr(o1), _ptr(o2), _ptr(o3), _ptr(i1), _ptr(i2), _ptr(i3), _ptr(h1), _ptr(h2),
                  C.byref(C.c_double(tolh)), _i(nmxhi))
    it = np.zeros(3, dtype=np.int32)
    check(lib().nekb_niterhm3(_ptr(it)))
    return [int(v) for v in it]


def set_projection(ifield: int, ifprojfld: bool, ldimt_proj: int = -1) -> None:
    check(lib().nekb_set_projection(ifield, int(ifprojfld), ldimt_proj))


def projection_reset() -> None:
    check(lib().nekb_projection_reset())


def hsolve(name: str, u, r, h1, h2, vmk, vml, imsh: int, tol: float, maxit: int, isd: int, approx, napprox, bi) -> int:
    """core/navier4.f:562 hsolve(name,u,r,h1,h2,vmk,vml,imsh,tol,maxit,isd,approx,napprox,bi); returns niterhm.
    napprox: int32 array (>= 2 entries) that receives ivar(1:2) = (mmx, m); approx is not used (device-resident space)."""
    nm = name.encode().ljust(4)[:4]
    lib().hsolve_(nm, _ptr(u), _ptr(r), _ptr(h1), _ptr(h2), _ptr(vmk), _ptr(vml), _i(imsh), C.byref(C.c_double(tol)), _i(maxit),
                  _i(isd), None if approx is None else _ptr(approx), _ptr(napprox), _ptr(bi), len(nm))
    return niterhm()


def set_mesh2(lx2: int, ixm12, dxm12, w3m2, metrics9, bm2=None, bm2inv=None, volvm2: float = 0.0, tolhs: float = 1e-8,
              nmxv: int = 1000, nelgv: int = 0, ifvcor: bool = False) -> None:
    """Mesh-2 (pressure, Gauss points) state of the Pn-Pn-2 operators.  ixm12, dxm12: (lx2, lx1) arrays as numpy sees the
    Fortran arrays (element [a, i] = ixm12(a,i)); metrics9: rxm2, sxm2, txm2, rym2, sym2, tym2, rzm2, szm2, tzm2."""
    f = lambda a: np.ascontiguousarray(np.asarray(a, dtype=np.float64).T).reshape(-1)     # -> column-major bytes
    I, D = f(ixm12), f(dxm12)
    w = np.ascontiguousarray(w3m2, dtype=np.float64).reshape(-1)
    mets = [np.ascontiguousarray(m, dtype=np.float64).reshape(-1) for m in metrics9]
    arr = (C.c_void_p * 9)(*[m.ctypes.data for m in mets])
    b = None if bm2 is None else np.ascontiguousarray(bm2, dtype=np.float64).reshape(-1)
    bi = None if bm2inv is None else np.ascontiguousarray(bm2inv, dtype=np.float64).reshape(-1)
    check(lib().nekb_set_mesh2(lx2, _ptr(I), _ptr(D), _ptr(w), C.cast(arr, C.c_void_p), None if b is None else _ptr(b),
                               None if bi is None else _ptr(bi), volvm2, tolhs, nmxv, nelgv, int(ifvcor)))


def opgradt(outx, outy, outz, inpfld) -> None:
    """core/navier1.f:4095 opgradt."""
    lib().opgradt_(_ptr(outx), _ptr(outy), _ptr(outz), _ptr(inpfld))


def opdiv(outfld, inpx, inpy, inpz) -> None:
    """core/navier1.f:4064 opdiv."""
    lib().opdiv_(_ptr(outfld), _ptr(inpx), _ptr(inpy), _ptr(inpz))


def opbinv(out1, out2, out3, inp1, inp2, inp3, h2inv) -> None:
    """core/navier1.f:775 opbinv."""
    lib().opbinv_(_ptr(out1), _ptr(out2), _ptr(out3), _ptr(inp1), _ptr(inp2), _ptr(inp3), _ptr(h2inv))


def cdabdtp(ap, wp, h1, h2, h2inv, intype: int) -> None:
    """core/navier1.f:258 cdabdtp."""
    lib().cdabdtp_(_ptr(ap), _ptr(wp), _ptr(h1), _ptr(h2), _ptr(h2inv), _i(intype))


def set_uzawa_state(tolps: float, param21: float = 0.0, prelax: float = 0.0, tolpdf: float = 0.0) -> None:
    check(lib().nekb_set_uzawa_state(tolps, param21, prelax, tolpdf))


def uzawa_gmres(res, h1, h2, h2inv, intype: int) -> int:
    """core/gmres.f:2 uzawa_gmres(res,h1,h2,h2inv,intype,iter); res is overwritten; returns iter."""
    it = C.c_int(0)
    lib().uzawa_gmres_(_ptr(res), _ptr(h1), _ptr(h2), _ptr(h2inv), _i(intype), C.byref(it))
    return int(it.value)


def set_param(idx: int, value: float) -> None:
    check(lib().nekb_set_param(idx, float(value)))


def set_binv(binvm1, bintm1=None) -> None:
    a = np.ascontiguousarray(binvm1, dtype=np.float64)
    b = None if bintm1 is None else np.ascontiguousarray(bintm1, dtype=np.float64)
    check(lib().nekb_set_binv(_ptr(a), None if b is None else _ptr(b)))


def cggos(u1, rhs1, x1, rmult, binv, tin: float, maxit: int, bpname: str = "bp5") -> int:
    m = C.c_int(maxit)
    lib().cggos_(_ptr(u1), _ptr(rhs1), _ptr(x1), _ptr(rmult), _ptr(binv), C.byref(C.c_double(tin)), C.byref(m),
                 bpname.encode(), len(bpname))
    return int(m.value)


def axhm1(ap1, p1, h1, h2, bpname: str = "bp5") -> float:
    pap = C.c_double(0.0)
    lib().axhm1_(C.byref(pap), _ptr(ap1), _ptr(p1), _ptr(h1), _ptr(h2), bpname.encode(), len(bpname))
    return float(pap.value)


def glsc3(a, b, mult) -> float:
    return float(lib().glsc3_(_ptr(a), _ptr(b), _ptr(mult), _i(len(a))))


# --------------------------------------------------------------------------------------------- pressure preconditioner
def h1mg_setup(fbc, xm1, ym1, zm1, vertex, nelv: int, null_space: bool = False) -> None:
    """core/hsmg.f:2234 h1mg_setup (+ swap_lengths, set_up_h1_crs) from the state the reference reads from COMMON:
    fbc[nelv,6] = get_fast_bc codes, /gxyz/ coordinates, /ivrtx/ vertex."""
    f = np.ascontiguousarray(fbc, dtype=np.int32).reshape(-1)
    v = np.ascontiguousarray(vertex, dtype=np.int64).reshape(-1)
    check(lib().nekb_h1mg_setup(f, *[np.ascontiguousarray(a, dtype=np.float64) for a in (xm1, ym1, zm1)], v, nelv, int(null_space)))


def h1mg_solve(z: np.ndarray, rhs: np.ndarray, if_hybrid: bool = False) -> None:
    """core/hsmg.f:1855 h1mg_solve(z,rhs,if_hybrid); rhs is masked in place as in the reference."""
    lib().h1mg_solve_(_ptr(z), _ptr(rhs), _i(1 if if_hybrid else 0))


def h1mg_info():
    lmax, crs = C.c_int(0), C.c_int(0)
    nh, nt = (C.c_int * 3)(), (C.c_int * 3)()
    check(lib().nekb_h1mg_info(C.byref(lmax), nh, nt, C.byref(crs)))
    return {"lmax": lmax.value, "nh": list(nh)[:lmax.value], "ntab": list(nt)[:lmax.value], "crs_iters": crs.value}


def h1mg_get(which: str, level: int, n: int) -> np.ndarray:
    out = np.zeros(n)
    check(lib().nekb_h1mg_get(which.encode(), level, _ptr(out), n))
    return out


def h1mg_free() -> None:
    lib().nekb_h1mg_free()


def hsmg_setup(fbc, xm1, ym1, zm1, vertex, nelv: int, null_space: bool, nelgv: int, df=None, sr=None, ss=None, st=None) -> None:
    """core/hsmg.f:22 hsmg_setup for the Pn-Pn-2 splitting; df, sr, ss, st are common /fastd/ as gen_fast leaves them
    (df(lx1^3,nelv); s?(2*lx1^2,nelv), S in the first half, column-major), or all None: gen_fast (core/fast3d.f:2-140) is
    run by the library."""
    f = np.ascontiguousarray(fbc, dtype=np.int32).reshape(-1)
    v = np.ascontiguousarray(vertex, dtype=np.int64).reshape(-1)
    a = [np.ascontiguousarray(q, dtype=np.float64).reshape(-1) for q in (xm1, ym1, zm1)]
    keep = [None if q is None else np.ascontiguousarray(q, dtype=np.float64).reshape(-1) for q in (df, sr, ss, st)]
    b = [None if q is None else _ptr(q) for q in keep]
    check(lib().nekb_hsmg_setup(f, *a, v, nelv, int(null_space), nelgv, *b))


def hsmg_solve(e: np.ndarray, r: np.ndarray) -> None:
    """core/hsmg.f:1376 hsmg_solve(e,r) on the (lx1-2)^3 pressure grid."""
    lib().hsmg_solve_(_ptr(e), _ptr(r))


def local_solves_fdm(u: np.ndarray, v: np.ndarray) -> None:
    """core/fasts.f:2 local_solves_fdm(u,v)."""
    lib().local_solves_fdm_(_ptr(u), _ptr(v))


def hsmg_get(which: str, level: int, n: int) -> np.ndarray:
    out = np.zeros(n)
    check(lib().nekb_hsmg_get(which.encode(), level, _ptr(out), n))
    return out


def fdm_h1_setup(face_internal, mask, xm1, ym1, zm1, nel: int) -> None:
    """core/hmholtz.f:1028-1220 set_fdm_prec_h1A for one field."""
    f = np.ascontiguousarray(face_internal, dtype=np.int32).reshape(-1)
    check(lib().nekb_fdm_h1_setup(f, *[np.ascontiguousarray(a, dtype=np.float64) for a in (mask, xm1, ym1, zm1)], nel))


def set_kfldfdm(kfldfdm: int) -> None:
    check(lib().nekb_set_kfldfdm(kfldfdm))


def set_fdm_prec_h1b(d: np.ndarray, h1: np.ndarray, h2: np.ndarray, nel: int) -> None:
    lib().set_fdm_prec_h1b_(_ptr(d), _ptr(h1), _ptr(h2), _i(nel))


def fdm_h1(z, r, d, mask, mult, nel: int, kt=None, rr=None) -> None:
    """core/hmholtz.f:937 fdm_h1(z,r,d,mask,mult,nel,kt,rr)."""
    lib().fdm_h1_(_ptr(z), _ptr(r), _ptr(d), _ptr(mask), _ptr(mult), _i(nel), None, None if rr is None else _ptr(rr))


def fdm_h1_get(which: str, nel: int):
    lx1 = _state["lx1"]
    out = {"ktype": np.zeros(3 * nel, dtype=np.int32), "elsize": np.zeros(3 * nel), "dd": np.zeros(9 * lx1)}[which]
    check(lib().nekb_fdm_h1_get(which.encode(), _ptr(out), out.nbytes))
    return out


def set_pressure_state(pmask, binvm1, tolps: float, param21: float = 0.0, ifvcor: bool = False, nelgv: int = 0) -> None:
    pm = np.ascontiguousarray(pmask, dtype=np.float64)
    bi = np.ascontiguousarray(binvm1, dtype=np.float64)
    check(lib().nekb_set_pressure_state(_ptr(pm), _ptr(bi), tolps, param21, int(ifvcor), nelgv))


def hmh_gmres(res: np.ndarray, h1: np.ndarray, h2: np.ndarray, wt: np.ndarray, maxit: int) -> int:
    """core/gmres.f:304 hmh_gmres(res,h1,h2,wt,iter): res is overwritten with the solution; returns iter."""
    it = C.c_int(maxit)
    lib().hmh_gmres_(_ptr(res), _ptr(h1), _ptr(h2), _ptr(wt), C.byref(it))
    return int(it.value)


def hmh_flex_cg(res: np.ndarray, h1: np.ndarray, h2: np.ndarray, wt: np.ndarray, maxit: int) -> int:
    """core/hmholtz.f:2164 hmh_flex_cg(res,h1,h2,wt,iter): res is overwritten with the solution; returns iter."""
    it = C.c_int(maxit)
    lib().hmh_flex_cg_(_ptr(res), _ptr(h1), _ptr(h2), _ptr(wt), C.byref(it))
    return int(it.value)


# --------------------------------------------------------------------------------------------- coarse-solver facade
def crs_setup(ids, Ai, Aj, A, null_space: bool, sid: int = 0) -> int:
    """core/fcrs.c:45 crs_setup(handle,sid,comm,np,n,id,nz,Ai,Aj,A,null_space,param,datafname,ierr) as called at
    core/navier8.f:217; returns the handle."""
    ids = np.ascontiguousarray(ids, dtype=np.int64).reshape(-1)
    Ai, Aj = (np.ascontiguousarray(a, dtype=np.int32).reshape(-1) for a in (Ai, Aj))
    A = np.ascontiguousarray(A, dtype=np.float64).reshape(-1)
    h, ierr, zero = C.c_int(-1), C.c_int(0), C.c_int(0)
    lib().crs_setup_(C.byref(h), _i(sid), C.byref(zero), _i(1), _i(ids.size), _ptr(ids), _i(A.size), _ptr(Ai), _ptr(Aj), _ptr(A),
                     _i(int(bool(null_space))), None, b"", C.byref(ierr))
    if ierr.value != 0 or h.value < 0:
        raise NekbError(lib().nekb_last_error().decode())
    return int(h.value)


def crs_solve(handle: int, b) -> np.ndarray:
    """core/fcrs.c:80 crs_solve(handle,x,b): x = Q A^-1 Q^T b on the handle's local dofs."""
    b = np.ascontiguousarray(b, dtype=np.float64).reshape(-1)
    x = np.zeros_like(b)
    lib().crs_solve_(_i(handle), _ptr(x), _ptr(b))
    return x


def crs_free(handle: int) -> None:
    lib().crs_free_(_i(handle))


def crs_amg_build_host(n: int, I, J, V, nmax: int = 4096, theta: float = 0.02, omega_p: float = 0.0):
    """Host set-up of the aggregation hierarchy (csrc/crs_amg.cuh).  Returns a list of levels, each a dict with `rowptr`, `col`,
    `val` (CSR) and, except for the coarsest, `agg` (row -> aggregate = row of the next level) and the prolongation `p_rowptr`,
    `p_col`, `p_val` (CSR, n_level x n_next; omega_p > 0: smoothed aggregation)."""
    I, J = (np.ascontiguousarray(a, dtype=np.int64).reshape(-1) for a in (I, J))
    V = np.ascontiguousarray(V, dtype=np.float64).reshape(-1)
    nl = C.c_int(0)
    check(lib().nekb_crs_amg_build_host(int(n), V.size, _ptr(I), _ptr(J), _ptr(V), int(nmax), float(theta), float(omega_p), C.byref(nl)))
    levels = []
    for l in range(nl.value):
        nn, nz = C.c_int64(0), C.c_int64(0)
        check(lib().nekb_crs_amg_level_info(l, C.byref(nn), C.byref(nz)))
        lv = dict(n=int(nn.value), rowptr=np.zeros(nn.value + 1, dtype=np.int64), col=np.zeros(nz.value, dtype=np.int32),
                  val=np.zeros(nz.value))
        agg = np.zeros(nn.value, dtype=np.int32) if l + 1 < nl.value else None
        check(lib().nekb_crs_amg_level_get(l, _ptr(lv["rowptr"]), _ptr(lv["col"]), _ptr(lv["val"]), None if agg is None else _ptr(agg)))
        if agg is not None:
            lv["agg"] = agg
            pz = C.c_int64(0)
            check(lib().nekb_crs_amg_level_p(l, C.byref(pz), None, None, None))
            lv["p_rowptr"], lv["p_col"], lv["p_val"] = np.zeros(nn.value + 1, dtype=np.int64), np.zeros(pz.value, dtype=np.int32), np.zeros(pz.value)
            check(lib().nekb_crs_amg_level_p(l, None, _ptr(lv["p_rowptr"]), _ptr(lv["p_col"]), _ptr(lv["p_val"])))
        levels.append(lv)
    return levels


# --------------------------------------------------------------------------------------------- device arrays
class DevArray:
    """A device buffer owned through the C-ABI helpers (section E of the header)."""

    def __init__(self, n: int, dtype=np.float64):
        self.n, self.dtype = int(n), np.dtype(dtype)
        self.ptr = lib().nekb_dev_alloc(max(self.n, 1) * self.dtype.itemsize)
        if not self.ptr:
            raise NekbError(lib().nekb_last_error().decode())

    @classmethod
    def from_host(cls, a: np.ndarray) -> "DevArray":
        a = np.ascontiguousarray(a)
        d = cls(a.size, a.dtype)
        check(lib().nekb_h2d(d.ptr, a.ctypes.data, a.nbytes))
        return d

    def to_host(self) -> np.ndarray:
        out = np.zeros(self.n, dtype=self.dtype)
        check(lib().nekb_d2h(out.ctypes.data, self.ptr, out.nbytes))
        return out

    def free(self) -> None:
        if self.ptr:
            lib().nekb_dev_free(self.ptr)
            self.ptr = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


# --------------------------------------------------------------------------------------------- multi-rank plumbing
def set_transport_torch(group=None) -> None:
    """Registers torch.distributed (any backend with CPU tensors, e.g. gloo) as the host transport of the
    setup code (numbering, shared-id discovery)."""
    import torch
    import torch.distributed as dist

    rank, world = dist.get_rank(group), dist.get_world_size(group)

    def _allgather(send, recv, nbytes, _user):
        try:
            src = torch.frombuffer((C.c_char * nbytes).from_address(send), dtype=torch.uint8).clone() if nbytes else torch.zeros(0, dtype=torch.uint8)
            outs = [torch.zeros(nbytes, dtype=torch.uint8) for _ in range(world)]
            dist.all_gather(outs, src, group=group)
            for r, t in enumerate(outs):
                if nbytes:
                    C.memmove(recv + r * nbytes, t.numpy().ctypes.data, nbytes)
            return 0
        except Exception as e:  # pragma: no cover
            print("transport allgather failed:", e)
            return 1

    def _alltoallv(send, send_bytes, recv, recv_bytes, _user):
        try:
            sb = [int(send_bytes[r]) for r in range(world)]
            rb = [int(recv_bytes[r]) for r in range(world)]
            so = np.concatenate([[0], np.cumsum(sb)]).astype(np.int64)
            ro = np.concatenate([[0], np.cumsum(rb)]).astype(np.int64)
            reqs, bufs = [], {}
            for r in range(world):
                if r == rank:
                    if sb[r]:
                        C.memmove(recv + int(ro[r]), send + int(so[r]), sb[r])
                    continue
                if rb[r]:
                    bufs[r] = torch.zeros(rb[r], dtype=torch.uint8)
                    reqs.append(dist.irecv(bufs[r], src=dist.get_global_rank(group, r) if group is not None else r, group=group))
                if sb[r]:
                    t = torch.frombuffer((C.c_char * sb[r]).from_address(send + int(so[r])), dtype=torch.uint8).clone()
                    reqs.append(dist.isend(t, dst=dist.get_global_rank(group, r) if group is not None else r, group=group))
            for q in reqs:
                q.wait()
            for r, t in bufs.items():
                C.memmove(recv + int(ro[r]), t.numpy().ctypes.data, rb[r])
            return 0
        except Exception as e:  # pragma: no cover
            print("transport alltoallv failed:", e)
            return 1

    ag, a2a = ALLGATHER_FN(_allgather), ALLTOALLV_FN(_alltoallv)
    _state["keep"] = [ag, a2a]  # keep the callbacks alive
    check(lib().nekb_set_transport(rank, world, ag, a2a, None))
    _state["np"] = world


def clear_transport() -> None:
    check(lib().nekb_set_transport(0, 1, ALLGATHER_FN(0), ALLTOALLV_FN(0), None))
    _state["np"] = 1
    _state["keep"] = []


def comm_init_torch(group=None) -> None:
    """Creates the library's NCCL communicator (one rank per GPU); the 128-byte unique id travels through
    torch.distributed."""
    import torch
    import torch.distributed as dist

    rank, world = dist.get_rank(group), dist.get_world_size(group)
    buf = (C.c_char * 128)()
    if rank == 0:
        check(lib().nekb_comm_unique_id(buf))
    backend = dist.get_backend(group)
    dev = torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")
    t = torch.frombuffer(buf, dtype=torch.uint8).clone().to(dev)
    dist.broadcast(t, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
    raw = bytes(t.cpu().numpy().tobytes())
    check(lib().nekb_comm_init(raw, rank, world))
    _state["np"] = world
