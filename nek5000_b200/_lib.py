"""ctypes binding of libnekb200.so (include/nekb200.h).  Loading fails loudly when the CUDA library has
not been built: there is no CPU fallback anywhere in this package."""
from __future__ import annotations

import ctypes as C
import os
import re

import numpy as np

from . import build as _build

HERE = os.path.dirname(os.path.abspath(__file__))
HEADER = os.path.join(HERE, "..", "include", "nekb200.h")


class NekbError(RuntimeError):
    pass


ALLGATHER_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p)
ALLTOALLV_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.POINTER(C.c_int64), C.c_void_p, C.POINTER(C.c_int64), C.c_void_p)

_lib = None

f64p = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
i64p = np.ctypeslib.ndpointer(dtype=np.int64, flags="C_CONTIGUOUS")
i32p = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")
vp = C.c_void_p
ip = C.POINTER(C.c_int)
dp = C.POINTER(C.c_double)


def declared_symbols() -> list[str]:
    """Every function name include/nekb200.h declares."""
    txt = open(HEADER).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    names = re.findall(r"\b([A-Za-z_][A-Za-z0-9_]*)\s*\(", txt)
    skip = {"defined", "C", "handler"}
    out = []
    for n in names:
        if n in skip or n.endswith("_fn"):
            continue
        if n.startswith("nekb_") or n.endswith("_"):
            if n not in out:
                out.append(n)
    return out


def lib() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    path = _build.LIB
    if not os.path.exists(path):
        raise NekbError(f"{path} is missing: build it with `python -m nek5000_b200.build` "
                        "(nvcc, sm_100a).  There is no CPU fallback.")
    L = C.CDLL(path, mode=C.RTLD_GLOBAL)
    L.nekb_last_error.restype = C.c_char_p
    L.nekb_stream.restype = vp
    L.nekb_launch_count.restype = C.c_int64
    L.nekb_launch_count.argtypes = [C.c_int]
    L.nekb_bp5_nel_local.restype = C.c_int64
    L.nekb_bp5_devptr.restype = vp
    L.nekb_bp5_devptr.argtypes = [C.c_char_p]
    L.nekb_dev_alloc.restype = vp
    L.nekb_dev_alloc.argtypes = [C.c_size_t]
    L.nekb_dev_free.argtypes = [vp]
    L.nekb_dev_free.restype = None
    L.nekb_h2d.argtypes = [vp, vp, C.c_size_t]
    L.nekb_d2h.argtypes = [vp, vp, C.c_size_t]
    L.nekb_init.argtypes = [C.c_int, C.c_int, C.c_int]
    L.nekb_finalize.restype = None
    L.nekb_prof_enable.argtypes = [C.c_int]
    L.nekb_prof_get.argtypes = [C.c_char_p, dp, C.POINTER(C.c_int64)]
    L.nekb_set_transport.argtypes = [C.c_int, C.c_int, ALLGATHER_FN, ALLTOALLV_FN, vp]
    L.nekb_comm_unique_id.argtypes = [vp]
    L.nekb_comm_init.argtypes = [vp, C.c_int, C.c_int]
    L.nekb_set_nel.argtypes = [C.c_int, C.c_int]
    L.nekb_set_gll.argtypes = [f64p, f64p]
    L.nekb_set_dxyz.argtypes = [f64p, f64p]
    L.nekb_set_geom.argtypes = [f64p] * 7
    L.nekb_set_geom_bp5.argtypes = [f64p]
    L.nekb_set_geom_from_xyz.argtypes = [f64p, f64p, f64p, C.c_int]
    L.nekb_get_geom.argtypes = [vp] * 8
    L.nekb_set_ifdfrm.argtypes = [vp]
    L.nekb_set_v1mask.argtypes = [f64p]
    L.nekb_set_ifield.argtypes = [C.c_int]
    L.nekb_set_field_handle.argtypes = [C.c_int, C.c_int]
    L.nekb_set_restol.argtypes = [C.c_int, C.c_double]
    L.nekb_last_history.argtypes = [vp, C.c_int64, ip, ip]
    L.nekb_ax_affine_deviation.restype = C.c_double
    L.nekb_d2d.argtypes = [vp, vp, C.c_size_t]
    L.nekb_set_step_info.argtypes = [C.c_int, C.c_double, C.c_double]
    L.nekb_gs_setup.argtypes = [ip, i64p, C.c_int64]
    L.nekb_gs_setup_dev.argtypes = [ip, vp, C.c_int64]
    L.nekb_gs_op_dev.argtypes = [C.c_int, vp, C.c_int, vp]
    L.nekb_gs_free.argtypes = [C.c_int]
    L.nekb_gs_info.argtypes = [C.c_int, C.POINTER(C.c_int64), C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
    L.nekb_gs_exchange_mode.argtypes = [C.c_int]
    L.nekb_fast1d_sem_host.argtypes = [C.c_int, C.c_int, C.c_int, C.c_double, C.c_double, C.c_double, vp, vp]
    L.nekb_fast1d_host.argtypes = [C.c_int, C.c_int, C.c_int, C.c_double, C.c_double, C.c_double, vp, vp]
    L.nekb_gs_get_map.argtypes = [C.c_int, i64p, i32p]
    L.nekb_gs_remote_info.argtypes = [C.c_int, ip, C.POINTER(C.c_int64)]
    L.nekb_gs_get_remote.argtypes = [C.c_int, i32p, i64p, i32p]
    L.nekb_ax_bp5_dev.argtypes = [vp, vp, vp]
    L.nekb_axhelm_dev.argtypes = [vp, vp, vp, vp, C.c_int]
    L.nekb_setprec_dev.argtypes = [vp, vp, vp, C.c_int]
    L.nekb_cggos_dev.argtypes = [vp, vp, vp, vp, C.c_double, C.c_int, ip, vp]
    L.nekb_cggo_dev.argtypes = [vp] * 7 + [C.c_int, C.c_double, C.c_int, ip, vp]
    L.nekb_setvert3d.argtypes = [i64p, C.POINTER(C.c_int64), C.c_int, C.c_int64, i64p, C.c_int]
    L.nekb_set_velocity_state.argtypes = [vp, vp, vp, vp]
    L.nekb_niterhm3.argtypes = [vp]
    L.nekb_ophinv_dev.argtypes = [vp] * 13 + [C.c_double, C.c_int, vp, vp]
    L.nekb_re2_info.argtypes = [C.c_char_p, C.POINTER(C.c_int64), ip, C.POINTER(C.c_int64), ip, C.POINTER(C.c_int64), ip, vp, C.c_int]
    L.nekb_re2_read_mesh.argtypes = [C.c_char_p, C.c_int64, C.c_int64, vp, vp, vp, vp]
    L.nekb_re2_read_bc.argtypes = [C.c_char_p, C.c_int, vp, vp]
    L.nekb_re2_read_curves.argtypes = [C.c_char_p, vp, vp]
    L.nekb_ma2_info.argtypes = [C.c_char_p, C.POINTER(C.c_int64), vp]
    L.nekb_ma2_read.argtypes = [C.c_char_p, C.c_int, C.c_int64, C.c_int64, vp, vp]
    L.crs_setup_.argtypes = [ip, ip, ip, ip, ip, vp, ip, vp, vp, vp, ip, vp, C.c_char_p, ip]
    L.crs_setup_.restype = None
    L.crs_solve_.argtypes = [ip, vp, vp]
    L.crs_solve_.restype = None
    L.crs_free_.argtypes = [ip]
    L.crs_free_.restype = None
    L.nekb_fcrs_solve_dev.argtypes = [C.c_int, vp, vp]
    L.nekb_crs_amg_build_host.argtypes = [C.c_int64, C.c_int64, vp, vp, vp, C.c_int64, C.c_double, C.c_double, ip]
    L.nekb_crs_amg_level_p.argtypes = [C.c_int, C.POINTER(C.c_int64), vp, vp, vp]
    L.nekb_crs_amg_level_info.argtypes = [C.c_int, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
    L.nekb_crs_amg_level_get.argtypes = [C.c_int, vp, vp, vp, vp]
    L.nekb_crs_amg_upload.argtypes = [C.c_double]
    L.nekb_crs_amg_solve_dev.argtypes = [vp, vp, C.c_double, C.c_int, ip]
    L.nekb_co2_info.argtypes = [C.c_char_p, C.POINTER(C.c_int64), C.POINTER(C.c_int64), ip]
    L.nekb_co2_read.argtypes = [C.c_char_p, C.c_int, C.c_int64, C.c_int64, vp, vp]
    L.nekb_assign_gllnid.argtypes = [vp, C.c_int64, C.c_int64, C.c_int]
    L.nekb_gs_discover.argtypes = [i64p, C.c_int64, ip, C.POINTER(C.c_int64), vp, vp, vp]
    L.nekb_bp5_setup.argtypes = [C.c_int] * 6 + [C.c_double]
    L.nekb_bp5_solve.argtypes = [C.c_double, C.c_int, ip, dp, vp]
    L.nekb_bp5_relerr.argtypes = [dp]
    L.nekb_bp5_get.argtypes = [C.c_char_p, vp, C.c_size_t]
    # Fortran-named entry points: everything by reference
    L.fgslib_gs_setup_.argtypes = [ip, i64p, ip, ip, ip]
    L.fgslib_gs_setup_.restype = None
    L.fgslib_gs_op_.argtypes = [ip, vp, ip, ip, ip]
    L.fgslib_gs_op_.restype = None
    L.fgslib_gs_op_many_.argtypes = [ip] + [vp] * 6 + [ip] * 4
    L.fgslib_gs_op_many_.restype = None
    L.fgslib_gs_op_fields_.argtypes = [ip, vp, ip, ip, ip, ip, ip]
    L.fgslib_gs_op_fields_.restype = None
    L.fgslib_gs_free_.argtypes = [ip]
    L.fgslib_gs_free_.restype = None
    L.setupds_.argtypes = [ip, ip, ip, ip, ip, ip, i64p, i64p]
    L.setupds_.restype = None
    L.dssum_.argtypes = [vp, ip, ip, ip]
    L.dssum_.restype = None
    L.dsop_.argtypes = [vp, C.c_char_p, ip, ip, ip, C.c_size_t]
    L.dsop_.restype = None
    L.axhelm_.argtypes = [vp, vp, vp, vp, ip, ip]
    L.axhelm_.restype = None
    L.setprec_.argtypes = [vp, vp, vp, ip, ip]
    L.setprec_.restype = None
    L.cggo_.argtypes = [vp, vp, vp, vp, vp, vp, ip, dp, ip, ip, vp, C.c_char_p, C.c_size_t]
    L.cggo_.restype = None
    L.cggos_.argtypes = [vp, vp, vp, vp, vp, dp, ip, C.c_char_p, C.c_size_t]
    L.cggos_.restype = None
    L.axhm1_.argtypes = [dp, vp, vp, vp, vp, C.c_char_p, C.c_size_t]
    L.axhm1_.restype = None
    L.glsc3_.argtypes = [vp, vp, vp, ip]
    L.glsc3_.restype = C.c_double
    L.hmholtz_.argtypes = [C.c_char_p, vp, vp, vp, vp, vp, vp, ip, dp, ip, ip, C.c_size_t]
    L.hmholtz_.restype = None
    L.vec_dssum_.argtypes = [vp, vp, vp, ip, ip, ip]
    L.vec_dssum_.restype = None
    L.vec_dsop_.argtypes = [vp, vp, vp, ip, ip, ip, C.c_char_p, C.c_size_t]
    L.vec_dsop_.restype = None
    L.nvec_dssum_.argtypes = [vp, ip, ip, ip]
    L.nvec_dssum_.restype = None
    L.dsavg_.argtypes = [vp]
    L.dsavg_.restype = None
    L.ophinv_.argtypes = [vp] * 8 + [dp, ip]
    L.ophinv_.restype = None
    L.hsolve_.argtypes = [C.c_char_p, vp, vp, vp, vp, vp, vp, ip, dp, ip, ip, vp, vp, vp, C.c_size_t]
    L.hsolve_.restype = None
    L.nekb_set_mesh2.argtypes = [C.c_int, vp, vp, vp, vp, vp, vp, C.c_double, C.c_double, C.c_int, C.c_int64, C.c_int]
    for nm, k in (("opgradt_", 4), ("opdiv_", 4), ("opbinv_", 7)):
        getattr(L, nm).argtypes = [vp] * k
        getattr(L, nm).restype = None
    L.cdabdtp_.argtypes = [vp] * 5 + [ip]
    L.cdabdtp_.restype = None
    L.hmh_flex_cg_.argtypes = [vp] * 4 + [ip]
    L.hmh_flex_cg_.restype = None
    L.uzawa_gmres_.argtypes = [vp] * 4 + [ip, ip]
    L.uzawa_gmres_.restype = None
    L.nekb_set_uzawa_state.argtypes = [C.c_double] * 4
    L.nekb_opgradt_dev.argtypes = [vp] * 4
    L.nekb_opdiv_dev.argtypes = [vp] * 4
    L.nekb_cdabdtp_dev.argtypes = [vp] * 5 + [C.c_int]
    L.nekb_uzawa_gmres_dev.argtypes = [vp] * 4 + [C.c_int, ip, dp, dp]
    L.nekb_set_projection.argtypes = [C.c_int, C.c_int, C.c_int]
    L.nekb_hsolve_dev.argtypes = [C.c_char_p] + [vp] * 6 + [C.c_int, C.c_double, C.c_int, vp, vp, ip]
    L.nekb_set_param.argtypes = [C.c_int, C.c_double]
    L.nekb_set_binv.argtypes = [vp, vp]
    # section F: pressure preconditioner + GMRES
    L.nekb_h1mg_setup.argtypes = [i32p, f64p, f64p, f64p, i64p, C.c_int, C.c_int]
    L.nekb_h1mg_solve_dev.argtypes = [vp, vp]
    L.h1mg_solve_.argtypes = [vp, vp, ip]
    L.h1mg_solve_.restype = None
    L.nekb_h1mg_schwarz_dev.argtypes = [C.c_int, vp, vp]
    L.nekb_crs_solve_dev.argtypes = [vp, vp]
    L.nekb_h1mg_info.argtypes = [ip, ip, ip, ip]
    L.nekb_h1mg_get.argtypes = [C.c_char_p, C.c_int, vp, C.c_size_t]
    L.nekb_crs_set_tolerance.argtypes = [C.c_double, C.c_int]
    L.nekb_h1mg_free.restype = None
    L.nekb_hsmg_setup.argtypes = [i32p, f64p, f64p, f64p, i64p, C.c_int, C.c_int, C.c_int64] + [C.c_void_p] * 4
    L.nekb_hsmg_solve_dev.argtypes = [vp, vp]
    L.nekb_local_solves_fdm_dev.argtypes = [vp, vp]
    L.hsmg_solve_.argtypes = [vp, vp]
    L.hsmg_solve_.restype = None
    L.local_solves_fdm_.argtypes = [vp, vp]
    L.local_solves_fdm_.restype = None
    L.nekb_hsmg_get.argtypes = [C.c_char_p, C.c_int, vp, C.c_size_t]
    L.nekb_fdm_h1_setup.argtypes = [i32p, f64p, f64p, f64p, f64p, C.c_int]
    L.nekb_set_kfldfdm.argtypes = [C.c_int]
    L.nekb_set_fdm_prec_h1b_dev.argtypes = [vp, vp, vp]
    L.nekb_fdm_h1_dev.argtypes = [vp, vp, vp, vp]
    L.set_fdm_prec_h1b_.argtypes = [vp, vp, vp, ip]
    L.set_fdm_prec_h1b_.restype = None
    L.fdm_h1_.argtypes = [vp, vp, vp, vp, vp, ip, vp, vp]
    L.fdm_h1_.restype = None
    L.nekb_fdm_h1_get.argtypes = [C.c_char_p, vp, C.c_size_t]
    L.nekb_set_pressure_state.argtypes = [vp, vp, C.c_double, C.c_double, C.c_int, C.c_int64]
    L.hmh_gmres_.argtypes = [vp, vp, vp, vp, ip]
    L.hmh_gmres_.restype = None
    L.nekb_hmh_gmres_dev.argtypes = [vp, vp, vp, vp, vp, C.c_double, C.c_int, ip, vp, dp]
    _lib = L
    return L


def check(rc: int) -> None:
    if rc != 0:
        raise NekbError(lib().nekb_last_error().decode())
