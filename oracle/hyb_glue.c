/* hyb_glue.c -- TEST INFRASTRUCTURE ONLY: the Fortran-side glue of INTEGRATION.md, written in C because no Fortran
 * compiler exists here.  It is linked into the DROP-IN variant of the transpiled reference (oracle/ref_build.py --hybrid):
 * the reference's own hmholtz / chktcg1 / hmh_gmres / bp5 ... run unchanged, while axhelm, dssum, dsop, cggo, cggos, axhm1,
 * h1mg_solve and gslib's / crs' Fortran API are undefined in that library and resolve to nek5000_b200/libnekb200.so.
 *
 * What this file does is what `subroutine nekb_register` / `subroutine h1mg_setup` of INTEGRATION.md do: hand the COMMON-block
 * state the replaced routines read implicitly to the library.  COMMON storage is reached through nekhyb_commons.h, generated
 * from the same SIZE as the translation ("a small generator reads SIZE -> header").
 */
#include <stdio.h>
#include <stdlib.h>

#include "nekb200.h"
#include "nekhyb_commons.h"

extern void get_fast_bc_(int *lbr, int *rbr, int *lbs, int *rbs, int *lbt, int *rbt, int *e, int *bsym, int *ierr);
extern void geodatstd_(double *gf);

static void chk(int rc, const char *what)
{
  if (rc) {
    fprintf(stderr, "hyb_glue: %s failed: %s\n", what, nekb_last_error());
    abort();
  }
}

/* before the first gs_setup: the library needs its device and the element counts (nek_init: after initdim / readat) */
void nekhyb_init_(const int *device)
{
  chk(nekb_init(*device, NEKHYB_LX1, 3), "nekb_init");
  chk(nekb_set_nel(*V_nelv, *V_nelt), "nekb_set_nel");
}

/* wherever Nek changes ifield (and once after setupds): dssum reads gsh_fld(ifield) from /comm_handles/ */
void nekhyb_set_field_(void)
{
  const int ifield = *V_ifield;
  chk(nekb_set_ifield(ifield), "nekb_set_ifield");
  chk(nekb_set_field_handle(ifield, V_gsh_fld[ifield - NEKHYB_GSH_FLD_LOW]), "nekb_set_field_handle");
}

/* once per step (cheap scalars): istep, volumes, the param(*) entries the replaced routines read */
void nekhyb_step_(void)
{
  static const int idx[] = {18, 21, 22, 42, 93, 94, 95};
  chk(nekb_set_step_info(*V_istep, *V_volvm1, *V_voltm1), "nekb_set_step_info");
  for (unsigned k = 0; k < sizeof idx / sizeof idx[0]; k++) chk(nekb_set_param(idx[k], V_param[idx[k] - 1]), "nekb_set_param");
  nekhyb_set_field_();
}

/* once after gengeom / geom_reset (INTEGRATION.md `nekb_register`) */
void nekhyb_register_(void)
{
  static int idf[NEKHYB_LELT];
  chk(nekb_set_gll(V_zgm1, V_wxm1), "nekb_set_gll");               /* zgm1(1,1): first column = the r direction */
  chk(nekb_set_dxyz(V_dxm1, V_dxtm1), "nekb_set_dxyz");
  chk(nekb_set_geom(V_g1m1, V_g2m1, V_g3m1, V_g4m1, V_g5m1, V_g6m1, V_bm1), "nekb_set_geom");
  for (int e = 0; e < *V_nelt; e++) idf[e] = V_ifdfrm[e] != 0;
  chk(nekb_set_ifdfrm(idf), "nekb_set_ifdfrm");
  chk(nekb_set_binv(V_binvm1, V_bintm1), "nekb_set_binv");
  chk(nekb_set_velocity_state(V_v1mask, V_v2mask, V_v3mask, V_vmult), "nekb_set_velocity_state");
  nekhyb_step_();
}

/* overrides core/hsmg.f:2234 h1mg_setup (called by the reference's set_overlap, core/navier6.f:29-101) */
void h1mg_setup_(void)
{
  static int fbc[6 * NEKHYB_LELT];
  int two = 2, ierr = 0;
  for (int e = 1; e <= *V_nelv; e++) {                             /* core/fast3d.f:802 -- lbr,rbr,lbs,rbs,lbt,rbt */
    int *f = fbc + 6 * (e - 1);
    get_fast_bc_(f, f + 1, f + 2, f + 3, f + 4, f + 5, &e, &two, &ierr);
  }
  const int nullsp = *V_ifvcor != 0;
  chk(nekb_h1mg_setup(fbc, V_xm1, V_ym1, V_zm1, (const int64_t *)V_vertex, *V_nelv, nullsp), "nekb_h1mg_setup");
  chk(nekb_set_pressure_state(V_pmask, V_binvm1, *V_tolps, V_param[20], nullsp, (int64_t)*V_nelgv), "nekb_set_pressure_state");
}

/* examples/bp5: cggos_ / axhm1_ read gf (common /bpgfactors/) and v1mask.  geodatstd is the reference's own routine; the bp5
 * driver recomputes the same factors right afterwards (bp5.usr:351). */
void nekhyb_bp5_register_(double *gf)
{
  geodatstd_(gf);
  chk(nekb_set_geom_bp5(gf), "nekb_set_geom_bp5");
  chk(nekb_set_v1mask(V_v1mask), "nekb_set_v1mask");
  nekhyb_step_();
}

/* the replaced cggo leaves its iteration count in the library; the reference's callers read common /iterhm/ niterhm */
void nekhyb_fetch_niterhm_(void) { *V_niterhm = nekb_niterhm(); }
