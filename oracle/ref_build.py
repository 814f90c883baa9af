#!/usr/bin/env python
"""Builds oracle/_ref/libnekref_<tag>.so -- the REFERENCE's own routines for the hot path, executing here.

TEST INFRASTRUCTURE ONLY.  No Fortran compiler exists in this image, so the recipe is:

  1. write a case SIZE file (compile-time extents; same keys as core/SIZE.template) into oracle/_ref/<tag>/SIZE;
  2. translate the needed routines of /root/reference/core/*.f, /root/reference/3rd_party/blasLapack/*.f and
     /root/reference/examples/bp5/bp5.usr with oracle/f77c.py (sources are read where they lie, never copied) into
     oracle/_ref/nekref_<tag>.c, together with nekref_<tag>.json (COMMON-block map used by oracle/ref.py);
  3. gcc -O2 -ffp-contract=off (the reference's default -O2 on baseline x86-64: no FMA contraction) that C + oracle/ref_stubs.c
     (serial gslib / crs stand-ins) into oracle/_ref/libnekref_<tag>.so.

oracle/_ref/ is git-ignored and travels to the GPU box as a prebuilt file.  Usage:

    python oracle/ref_build.py [--lx1 8] [--lx2 8] [--lelt 64] [--tag lx8] [--force]
"""
from __future__ import annotations

import argparse
import glob
import json
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("NEK_REFERENCE", "/root/reference")
OUT = os.path.join(HERE, "_ref")
sys.path.insert(0, HERE)

CORE_FILES = """speclib.f mxm_std.f mxm_wrapper.f math.f hmholtz.f coef.f dssum.f navier8.f hsmg.f fast3d.f fasts.f gmres.f
navier1.f navier4.f navier5.f ic.f genxyz.f connect1.f connect2.f bdry.f subs1.f subs2.f comm_mpi.f mpi_dummy.f induct.f
navier2.f navier3.f navier6.f navier7.f eigsolv.f map2.f gauss.f drive2.f drive1.f calcz.f convect.f prepost.f plan4.f
plan5.f mvmesh.f conduct.f perturb.f vprops.f hpf.f ssolv.f planx.f postpro.f multimesh.f makeq.f makeq_aux.f
dprocmap.f interp.f gfldr.f hrefine.f pertsupport.f convect2.f navier0.f intp_usr.f reader_re2.f reader_rea.f
reader_par.f byte_mpi.f""".split()

# entry points the tests / golden generator / CPU baseline call; everything they call is followed automatically
ROOTS = """initdim initdat initds dsset setedge geom2 rone rzero copy invcol1 invcol2 zwgll zwgl dgll dgllgl igllm iglm genwz geom1 glmapm1 geodat1 volume setinvm setdef gencoor genxyz xyzlin
axhelm setfast sfastax setprec cggo hmholtz chktcg1 fdm_h1 set_fdm_prec_h1a set_fdm_prec_h1b gen_fast_spacing
dssum dsop vec_dssum setupds set_vert setvert3d get_vert dsavg bcmask setup_topo
glsc3 glsc2 glsum glamax glmax glmin vlsc3 vlsc32 add2s1 add2s2 col2 col3 mxm
bp5 sin_fld_h1 cggos ax_e_bp5 axhm1_bp5 geodatstd setprecn xmask1 glrdif rand_fld_h1 ran1 loc_grad3 loc_grad3t
h1mg_setup h1mg_solve hsmg_setup hsmg_solve hmh_gmres uzawa_gmres ax set_up_h1_crs crs_solve_h1 cdabdtp
local_solves_fdm gen_fast gen_fast_g semhat generalev hsmg_setup_semhat hsmg_setup_intp h1mg_setup_fdm h1mg_schwarz h1mg_rstr hsmg_intp
hsolve ophinv opgradt opdiv opbinv ortho project1 project2 hmhzpf set_overlap swap_lengths hsmg_index_0
assign_gllnid isort iswapt_ip""".split()

# never followed: file output, whole-application drivers, things a unit test never reaches
STOP = """outpost outpost2 outfld exitt exitti exittr printpartstat usrsetvert dnekclock dnekclock_sync nekgsync userchk
userbc userf userq useric uservp usrdat usrdat2 usrdat3 nek_init nek_solve nek_end nek_advance in_situ_check
mfo_open_files mfi fgslib_gs_unique plan4 plan3 fluid heat cvode_solve readat readat_par readat_big
full_restart_save restart hpts prepost lastep drgtrq""".split()


# routines whose write(6,...) statements report their numeric items through f77_trace (oracle/f77c.py trace_write):
# the residual histories the reference only logs
TRACE_UNITS = """cggo hmh_gmres hmh_flex_cg uzawa_gmres""".split()


# ---- the DROP-IN variant ("hyb"): the reference's own callers on top of libnekb200.so ------------------------------------
# The routines below are NOT translated; the library is linked against nek5000_b200/libnekb200.so, so the transpiled
# reference's calls to them (Fortran convention: by reference, hidden CHARACTER lengths) land in the CUDA entry points --
# what a Nek5000 build gets when libnekb200.so precedes libnek5000.a on the link line.  The gslib and crs
# stand-ins of ref_stubs.c are compiled out as well (fgslib_gs_* / crs_* come from the product), and oracle/hyb_glue.c
# plays the Fortran glue of INTEGRATION.md: it registers the COMMON state (located through a generated header of COMMON
# offsets, the "SIZE -> header generator" of SURVEY 8b) and overrides h1mg_setup.
HYB_STOP = """axhelm dssum dsop cggo cggos axhm1 h1mg_solve h1mg_setup""".split()
HYB_ROOTS = """get_fast_bc get_vert set_overlap""".split()
# COMMON variables the glue reads
HYB_VARS = """nelv nelt nelgv zgm1 wxm1 dxm1 dxtm1 g1m1 g2m1 g3m1 g4m1 g5m1 g6m1 bm1 binvm1 bintm1 ifdfrm istep volvm1 voltm1
ifield gsh_fld param v1mask v2mask v3mask vmult pmask tolps ifvcor xm1 ym1 zm1 vertex niterhm""".split()


def hyb_header(meta, lx1):
    """C view of the COMMON variables in HYB_VARS: extern storage of the block + a typed pointer macro per variable."""
    ct = {"r8": "double", "i4": "int", "l4": "int", "i8": "long long"}
    blocks, lines = set(), []
    for v in HYB_VARS:
        m = meta["commons"][v]
        blocks.add(m["block"])
        lines.append(f"#define V_{v} (({ct[m['type']]} *)(cb_{m['block']} + {m['offset']}))")
    lo = meta["commons"]["gsh_fld"]["lows"][0]
    head = ["/* generated by oracle/ref_build.py from the same SIZE the reference was translated with */",
            f"#define NEKHYB_LX1 {lx1}", f"#define NEKHYB_LELT {meta['lelt']}", f"#define NEKHYB_GSH_FLD_LOW {lo}"]
    head += [f"extern char cb_{b}[];" for b in sorted(blocks)]
    return "\n".join(head + lines) + "\n"


def size_text(lx1, lx2, lelt, lgmres, ldimt=1):
    lxd = (3 * lx1 + 1) // 2
    return f"""c     generated by oracle/ref_build.py (keys of core/SIZE.template)
      integer ldim,lx1,lxd,lx2,lx1m,lelg,lelt,lpmin,ldimt
      integer lpelt,lbelt,toteq,lcvelt
      integer lelx,lely,lelz,mxprev,lgmres,lorder,lhis
      integer maxobj,lpert,nsessmax,lxo
      integer lfdm,ldimt_proj,lelr
      parameter (ldim=3)
      parameter (lx1={lx1})
      parameter (lxd={lxd})
      parameter (lx2={lx2})
      parameter (lelg={lelt})
      parameter (lpmin=1)
      parameter (lelt={lelt})
      parameter (ldimt={ldimt})
      parameter (ldimt_proj=1)
      parameter (lelr=lelt)
      parameter (lhis=1)
      parameter (maxobj=1)
      parameter (lpert=1)
      parameter (toteq=1)
      parameter (nsessmax=1)
      parameter (lxo=lx1)
      parameter (mxprev=20,lgmres={lgmres})
      parameter (lorder=3)
      parameter (lx1m=1)
      parameter (lfdm=0)
      parameter (lelx=1,lely=1,lelz=1)
      parameter (lbelt=1)
      parameter (lpelt=1)
      parameter (lcvelt=1)
      include 'SIZE.inc'
"""


def build(lx1=8, lx2=None, lelt=64, lgmres=30, tag=None, force=False, verbose=False, hybrid=False):
    import f77c
    lx2 = lx1 if lx2 is None else lx2
    tag = tag or f"lx{lx1}" + ("" if lx2 == lx1 else f"p{lx2}") + f"e{lelt}" + ("" if lgmres == 30 else f"g{lgmres}") + \
        ("hyb" if hybrid else "")
    so = os.path.join(OUT, f"libnekref_{tag}.so")
    deps = [os.path.join(HERE, f) for f in ("f77c.py", "ref_build.py", "ref_stubs.c") + (("hyb_glue.c",) if hybrid else ())]
    if os.path.exists(so) and not force and all(os.path.getmtime(so) >= os.path.getmtime(d) for d in deps):
        return so
    if not os.path.isdir(os.path.join(REF, "core")):
        if os.path.exists(so):
            return so          # prebuilt (GPU box): the reference tree does not travel
        raise FileNotFoundError(f"{REF} is absent and {so} was not prebuilt")
    inc = os.path.join(OUT, tag)
    os.makedirs(inc, exist_ok=True)
    with open(os.path.join(inc, "SIZE"), "w") as f:
        f.write(size_text(lx1, lx2, lelt, lgmres))
    tr = f77c.Translator(REF, [inc, os.path.join(REF, "core")],
                         {"PARALLEL": "PARALLEL.default", "mpif.h": "mpi_dummy.h"}, defines=["UNDERSCORE"])
    tr.trace_units = set(TRACE_UNITS)
    stop = STOP + (HYB_STOP if hybrid else [])
    for fn in CORE_FILES:
        tr.add_file(os.path.join(REF, "core", fn), skip=stop)
    tr.add_file(os.path.join(REF, "examples", "bp5", "bp5.usr"), skip=stop)
    for fn in sorted(glob.glob(os.path.join(REF, "3rd_party", "blasLapack", "*.f"))):
        tr.add_file(fn)
    roots = [r for r in ROOTS + (HYB_ROOTS if hybrid else []) if r in tr.units]
    code, missing = tr.translate(roots, stop_at=set(stop))
    csrc = os.path.join(OUT, f"nekref_{tag}.c")
    with open(csrc, "w") as f:
        f.write(code)
    meta = dict(lx1=lx1, lx2=lx2, lelt=lelt, lgmres=lgmres, tag=tag, commons=tr.common_map(), commons_alt=tr.common_alt(),
                common_size=tr.common_size, failed=tr.failed, missing=missing,
                absent_roots=[r for r in ROOTS if r not in tr.units])
    with open(os.path.join(OUT, f"nekref_{tag}.json"), "w") as f:
        json.dump(meta, f)
    cc = "/usr/bin/gcc" if os.path.exists("/usr/bin/gcc") else "gcc"
    # static COMMON storage is ~2.4 MB per element of lelt: beyond 2 GB the data needs the medium code model
    big = ["-mcmodel=medium"] if sum(tr.common_size.values()) > (1 << 30) else []
    # -Bsymbolic: the library's references to its own globals (fgslib_gs_setup_, axhelm_, cggo_, ... -- the very names the
    # product exports for the drop-in) must bind inside the library even when libnekb200.so is loaded in the same process
    cmd = [cc, "-O2", "-ffp-contract=off", "-w", "-fPIC", *big, "-shared", "-Wl,-Bsymbolic", "-o", so, csrc,
           os.path.join(HERE, "ref_stubs.c"), "-lm"]
    if hybrid:
        # -Bsymbolic stays: the routines this library DEFINES (hmholtz_, hmh_gmres_, bp5_, glsc3_, ...) keep calling each
        # other, so exactly the HYB_STOP routines (undefined here) and the gslib / crs API resolve to libnekb200.so.
        with open(os.path.join(inc, "nekhyb_commons.h"), "w") as f:
            f.write(hyb_header(meta, lx1))
        prod = os.path.normpath(os.path.join(HERE, "..", "nek5000_b200"))
        cmd = cmd[:-1] + ["-DNEKHYB", os.path.join(HERE, "hyb_glue.c"), "-I", inc, "-I", os.path.join(HERE, "..", "include"),
                          "-L", prod, "-lnekb200", "-Wl,-rpath,$ORIGIN/../../nek5000_b200", "-Wl,--no-undefined", "-lm"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("gcc failed on the transpiled reference:\n" + r.stderr[-4000:])
    if verbose:
        print(f"built {so}: {code.count(chr(10))} lines of C, {len(tr.failed)} untranslatable units (abort stubs): "
              f"{sorted(tr.failed)}; unresolved externals: {missing}")
    return so


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--lx1", type=int, default=8)
    ap.add_argument("--lx2", type=int, default=None)
    ap.add_argument("--lelt", type=int, default=64)
    ap.add_argument("--lgmres", type=int, default=30)
    ap.add_argument("--tag", default=None)
    ap.add_argument("--force", action="store_true")
    ap.add_argument("--hybrid", action="store_true", help="the drop-in variant: reference callers on top of libnekb200.so")
    a = ap.parse_args()
    print(build(a.lx1, a.lx2, a.lelt, a.lgmres, a.tag, a.force, verbose=True, hybrid=a.hybrid))
