/*
 * nekb200.h -- C-ABI of libnekb200.so: the B200-native (sm_100a, FP64 CUDA) replacement for
 * Nek5000's Helmholtz / gather-scatter / PCG hot path.
 *
 * Every entry point names the reference interface it replaces (paths relative to the
 * Nek5000 tree).  Plain pointers and sizes only; no torch / C++ types.
 *
 * There is no plugin registry in Nek5000: a replacement overrides by SYMBOL NAME at link
 * time (core/makenek.inc:304-307 adds --allow-multiple-definition; bin/makenek:31-34 lets a
 * case add objects).  Section A therefore exports the gfortran-mangled names
 * (lower case + trailing underscore, core/name.h:32-41) with Fortran by-reference
 * arguments.  Because most operands of those routines are COMMON-block state and not
 * arguments (SURVEY.md 8b "implicit state"), section B is the explicit registration the
 * Fortran side performs once after gengeom.  Section C is the same
 * functionality on device-resident buffers (used by the BP5 driver and by callers that keep
 * a whole solve on the GPU).  Section D is the host-side setup that feeds the path
 * (numbering, geometry, synthetic BP5 case).
 *
 * Error convention (reference: print on rank 0 + call exitt, core/comm_mpi.f:550-636):
 * nekb_* functions return 0 on success and non-zero on failure with the message
 * retrievable through nekb_last_error(); the Fortran-named entry points have no return
 * value, so on failure they print "nekb200: <file:line> <message>" to stderr and call the
 * registered exit handler (default: abort()), never returning stale output.
 *
 * Threading: one host thread per process, one GPU per process (Nek's rank model); all GPU
 * work is enqueued on the library's stream; host-buffer entry points are synchronous.
 */
#ifndef NEKB200_H
#define NEKB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ------------------------------------------------------------------------------------
 * Lifecycle
 * ---------------------------------------------------------------------------------- */

/* Selects the CUDA device, creates the stream and scratch.  lx1 = GLL points per direction
 * (core/SIZE.template:13 lx1); ldim must be 3.  Idempotent for identical arguments. */
int nekb_init(int device, int lx1, int ldim);
void nekb_finalize(void);
const char *nekb_last_error(void);
/* Handler called by the Fortran-named entry points on failure (reference: exitt,
 * core/comm_mpi.f:550).  NULL restores abort(). */
void nekb_set_exit_handler(void (*handler)(void));
/* Library stream as a cudaStream_t cast to void* (so callers can order their own work). */
void *nekb_stream(void);
/* Number of kernel launches issued by the library since the last reset (bench.py's
 * gpu_launches claim). */
int64_t nekb_launch_count(int reset);

/* Per-kernel device time of the cggos loop, measured with CUDA events on the library stream while enabled
 * (reference analogue: the taxhm/tdsum timers, core/hmholtz.f:113,257, core/dssum.f).  kernel: "ax" (Ax incl.
 * the fused pap), "gs" (gather-scatter incl. exchange), "update" (x,r update + weighted dot), "pupdate". */
int nekb_prof_enable(int on);
int nekb_prof_get(const char *kernel, double *seconds, int64_t *launches);

/* ------------------------------------------------------------------------------------
 * Multi-rank transport (host side, setup only) and device collectives
 * ---------------------------------------------------------------------------------- */

/* Host collectives the setup code needs when np > 1 (reference: gslib crystal router /
 * MPI inside fgslib_gs_setup and gbtuple_rank8, core/navier8.f:1934-2002).  Supplied by the
 * host program: MPI in a Fortran build, torch.distributed (gloo) in the Python harness.
 * Counts are in bytes.  Both return 0 on success. */
typedef int (*nekb_allgather_fn)(const void *send, void *recv, size_t bytes_per_rank, void *user);
typedef int (*nekb_alltoallv_fn)(const void *send, const int64_t *send_bytes, void *recv,
                                 const int64_t *recv_bytes, void *user);
int nekb_set_transport(int rank, int nranks, nekb_allgather_fn allgather, nekb_alltoallv_fn alltoallv,
                       void *user);

/* Device collectives over NCCL (NVLink 5 / NVSwitch); replaces gop -> mpi_allreduce
 * (core/comm_mpi.f:216-259) and gslib's pairwise exchange.  id_out/id_in: 128-byte
 * ncclUniqueId produced on rank 0 and distributed by the host program. */
int nekb_comm_unique_id(void *id_out_128);
int nekb_comm_init(const void *id_in_128, int rank, int nranks);

/* ------------------------------------------------------------------------------------
 * A. Fortran-named drop-in entry points (host buffers, synchronous)
 * ---------------------------------------------------------------------------------- */

/* gslib v1.0.9 Fortran API as called at core/dssum.f:20,79,198,277 and core/hsmg.f:335-362.
 * dom: 1 = double (only supported datatype here); op: 1 +, 2 *, 3 min, 4 max; transpose 0. */
void fgslib_gs_setup_(int *handle, const int64_t *id, const int *n, const int *comm, const int *np);
void fgslib_gs_op_(const int *handle, double *u, const int *dom, const int *op, const int *transpose);
void fgslib_gs_op_many_(const int *handle, double *u1, double *u2, double *u3, double *u4, double *u5,
                        double *u6, const int *n, const int *dom, const int *op, const int *transpose);
void fgslib_gs_op_fields_(const int *handle, double *u, const int *stride, const int *n, const int *dom,
                          const int *op, const int *transpose);
void fgslib_gs_free_(const int *handle);

/* core/dssum.f:1 setupds(gs_handle,nx,ny,nz,nel,melg,vertex,glo_num): set_vert
 * (core/navier8.f:3-33 -> setvert3d :2004-2360) then fgslib_gs_setup. */
void setupds_(int *gs_handle, const int *nx, const int *ny, const int *nz, const int *nel,
              const int *melg, const int64_t *vertex, int64_t *glo_num);
/* core/dssum.f:33 dssum(u,nx,ny,nz): gs_op(+) on gsh_fld(ifield) (see nekb_set_ifield). */
void dssum_(double *u, const int *nx, const int *ny, const int *nz);
/* core/dssum.f:100 dsop(u,op,nx,ny,nz), op is character*3: '+  ','sum','*  ','mul','m  ','min','mna',
 * 'M  ','max','mxa' (core/dssum.f:110-158); trailing hidden length as gfortran passes it. */
void dsop_(double *u, const char *op, const int *nx, const int *ny, const int *nz, size_t op_len);
/* core/dssum.f:163 vec_dssum(u,v,w,nx,ny,nz), :198 vec_dsop(u,v,w,nx,ny,nz,op), :260 nvec_dssum(u,stride,n,gs_handle);
 * core/ic.f:1871 dsavg(u) = vmult * dssum(u) (vmult from nekb_set_velocity_state). */
void vec_dssum_(double *u, double *v, double *w, const int *nx, const int *ny, const int *nz);
void vec_dsop_(double *u, double *v, double *w, const int *nx, const int *ny, const int *nz, const char *op, size_t op_len);
void nvec_dssum_(double *u, const int *stride, const int *n, const int *gs_handle);
void dsavg_(double *u);

/* core/hmholtz.f:72 axhelm(au,u,helm1,helm2,imesh,isd) -- general 3-D branch :191-217,:225. */
void axhelm_(double *au, const double *u, const double *helm1, const double *helm2, const int *imesh,
             const int *isd);
/* core/hmholtz.f:380 setprec(dpcm1,helm1,helm2,imsh,isd) incl. dssum + invcol1 (:520-521). */
void setprec_(double *dpcm1, const double *helm1, const double *helm2, const int *imsh, const int *isd);
/* core/hmholtz.f:611 cggo(x,f,h1,h2,mask,mult,imsh,tin,maxit,isd,binv,name): Jacobi-PCG branch
 * (kfldfdm<0) and Schwarz branch (kfldfdm>=0, fdm_h1, :737-745).  name is character*4 with gfortran's
 * hidden trailing length.  name = 'PRES' follows param(42) as the reference does (:641-657): 0 ->
 * hmh_gmres, 2 -> hmh_flex_cg, 1 -> this routine's own PCG with the 'PRES' extras (:710-712,
 * :741-748: coarse-grid correction crs_solve_h1 on the coarse solver of nekb_h1mg_setup, ortho in
 * every iteration, tol = |param(21)| when non-zero).  The iteration count is left in
 * nekb_niterhm() (reference: common /iterhm/ niterhm, :638). */
void cggo_(double *x, const double *f, const double *h1, const double *h2, const double *mask,
           const double *mult, const int *imsh, const double *tin, const int *maxit, const int *isd,
           const double *binv, const char *name, size_t name_len);
/* examples/bp5/bp5.usr:797 cggos(u1,rhs1,x1,rmult,binv,tin,maxit,bpname) with bpname='bp5'.
 * On return *maxit holds the iterations performed (bp5.usr:893). */
void cggos_(double *u1, const double *rhs1, const double *x1, const double *rmult, const double *binv,
            const double *tin, int *maxit, const char *bpname, size_t bpname_len);
/* examples/bp5/bp5.usr:1389 axhm1(pap,ap1,p1,h1,h2,bpname) with bpname='bp5' (:1309-1341). */
void axhm1_(double *pap, double *ap1, const double *p1, const double *h1, const double *h2,
            const char *bpname, size_t bpname_len);
/* core/math.f:775 glsc3(a,b,mult,n) = global sum a*b*mult (gop '+'). */
double glsc3_(const double *a, const double *b, const double *mult, const int *n);
/* core/hmholtz.f:2 hmholtz(name,u,rhs,h1,h2,mask,mult,imsh,tli,maxit,isd): dssum + mask of rhs (in place, as the reference),
 * chktcg1 (:527-609) when param(22) = 0 or istep <= 10, then cggo with binvm1/bintm1.  name is character*4; gfortran
 * appends its hidden length after the last argument. */
void hmholtz_(const char *name, double *u, double *rhs, const double *h1, const double *h2, const double *mask,
              const double *mult, const int *imsh, const double *tli, const int *maxit, const int *isd, size_t name_len);
int nekb_niterhm(void);

/* ------------------------------------------------------------------------------------
 * B. Registration of the COMMON-block state the Fortran routines read implicitly
 * ---------------------------------------------------------------------------------- */

/* /dimn/ nelv, nelt (core/SIZE.inc) */
int nekb_set_nel(int nelv, int nelt);
/* /dxyz/ dxm1, dxtm1 (core/DXYZ:4-9); dym1/dzm1 are identical on the GLL mesh
 * (core/coef.f:271-273).  lx1*lx1 doubles each, Fortran order. */
int nekb_set_dxyz(const double *dxm1, const double *dxtm1);
/* /gauss/ zgm1(lx1,1), wxm1(lx1) (core/WZ): GLL points and weights.  Optional: without it (and without
 * nekb_set_dxyz) the library computes its own GLL operators (core/speclib.f:107 ZWGLL, :800 DGLL). */
int nekb_set_gll(const double *zgm1, const double *wxm1);
/* /gmfact/ g1m1..g6m1 (core/GEOM:42-48, order rr,ss,tt,rs,rt,st) and /mass/ bm1 (core/MASS:4),
 * each lx1^3*nelt doubles.  Call again whenever geom_reset/gengeom ran. */
int nekb_set_geom(const double *g1m1, const double *g2m1, const double *g3m1, const double *g4m1,
                  const double *g5m1, const double *g6m1, const double *bm1);
/* BP5's interleaved gf(6,lx1^3,nelt), order rr,rs,rt,ss,st,tt (examples/bp5/bp5.usr:332,623-699). */
int nekb_set_geom_bp5(const double *gf);
/* Computes the factors on the device from /gxyz/ xm1,ym1,zm1 (host, lx1^3*nelt each) instead of
 * uploading them: bp5_form != 0 -> geodatstd (examples/bp5/bp5.usr:623-699), else glmapm1 + geodat1
 * (core/coef.f:555-631, :633-784), which also yields bm1. */
int nekb_set_geom_from_xyz(const double *xm1, const double *ym1, const double *zm1, int bp5_form);
/* Host copies of the registered / computed factors (any pointer may be NULL): core order g1m1..g6m1, bm1, and
 * BP5's interleaved gf(6,lx1^3,nelt). */
int nekb_get_geom(double *g1m1, double *g2m1, double *g3m1, double *g4m1, double *g5m1, double *g6m1, double *bm1,
                  double *gf);
/* /fastmd/ ifdfrm(lelt) (core/hmholtz.f:89): NULL = all elements deformed (param(59)=1 default,
 * core/reader_par.f:78). */
int nekb_set_ifdfrm(const int *ifdfrm);
/* v1mask of bp5.usr:142-153 xmask1 (COMMON v1mask, core/SOLN). */
int nekb_set_v1mask(const double *v1mask);
/* TSTEP ifield and /comm_handles/ gsh_fld(ifield) (core/PARALLEL.default:26-27), istep, and the
 * scalars cggo reads: volvm1, voltm1 (core/MASS), param(18,22) stay at their defaults. */
int nekb_set_ifield(int ifield);
int nekb_set_field_handle(int ifield, int gs_handle);
/* 1 when the registered geometric factors are, on every element and to 1e-13, a per-element constant times w_i w_j w_k
 * (affine elements: every genbox brick) and the fused BP5 operator kernel therefore reads six constants per element instead
 * of six factors per node (csrc/ax.cuh ax_cg_affine_kernel); 0 otherwise or with NEKB_AX_AFFINE=0.  The arithmetic is that
 * of core/hmholtz.f:191-217 / bp5.usr:1278-1341 with G_ab(i,j,k) written as c_ab * w3(i,j,k). */
int nekb_ax_affine_active(void);
/* Largest relative deviation of a registered factor from (element constant) * w3 found by the last check: the rounding noise
 * of the numerical differentiation behind the factors (2e-12 on a 64^3 box) for affine elements, O(1) for deformed ones. */
double nekb_ax_affine_deviation(void);
/* Residual history of the most recent cggo (rows of 3: rtz1, rbn2, rho per executed check, core/hmholtz.f:754-802) or
 * hmh_gmres (rows of 1: rnorm per iteration, core/gmres.f:486-498) solve, whichever entry point ran it -- the
 * numbers the reference prints per iteration and keeps nowhere.  out may be NULL to query the shape. */
int nekb_last_history(double *out, int64_t capacity, int *rows, int *cols);
/* TSTEP restol(0:ldimt1) (set by core/reader_par.f from residualTol / param(22)): a non-zero restol(ifield) overrules the
 * tolerance cggo receives unless that one is negative (core/hmholtz.f:673-679).  0 (default) = use the caller's. */
int nekb_set_restol(int ifield, double restol);
int nekb_set_step_info(int istep, double volvm1, double voltm1);
/* core/induct.f:1022-1090 ophinv(o1,o2,o3,i1,i2,i3,h1,h2,tolh,nmxhi): o_k = (h1 A + h2 B)^-1 i_k for the three velocity
 * components -- in the reference three hsolve -> hmholtz -> cggo calls in a row (standard branch: ifstrs = .false., no
 * residual projection), here ONE fused 3-right-hand-side PCG (hcg.cuh: factors, h1, h2 B and the diagonal are streamed once
 * per element for all three components; each component keeps its own scalars, tolerance (chktcg1) and exit test, so
 * iteration counts equal those of the three separate solves).  i1..i3 return dssum'ed and masked as hmholtz leaves them.
 * COMMON state: v1mask, v2mask, v3mask, vmult (core/SOLN) via nekb_set_velocity_state; binvm1 via nekb_set_binv;
 * volvm1/istep via nekb_set_step_info; param(22).  nekb_niterhm3: niterhm of the three solves (the reference's /iterhm/
 * holds the last).  nekb_ophinv_dev: the same on device pointers with explicit masks/mult/binv; hist_host (may be NULL)
 * receives 3 rows of 3*(min(maxit,900)+2) doubles (rtz1, rbn2, rho per iteration). */
void ophinv_(double *o1, double *o2, double *o3, double *i1, double *i2, double *i3, const double *h1, const double *h2,
             const double *tolh, const int *nmxhi);
int nekb_set_velocity_state(const double *v1mask, const double *v2mask, const double *v3mask, const double *vmult);
int nekb_niterhm3(int *niter3);
int nekb_ophinv_dev(double *o1, double *o2, double *o3, double *i1, double *i2, double *i3, const double *h1, const double *h2,
                    const double *m1, const double *m2, const double *m3, const double *mult, const double *binv, double tolh,
                    int maxit, int *niter3, double *hist_host);

/* core/navier4.f:562-634 hsolve(name,u,r,h1,h2,vmk,vml,imsh,tol,maxit,isd,approx,napprox,bi): the standard branch is
 * hmholtz; with residual projection (ifprojfld(ifield) or name = 'PRES', param(93) > 0, param(94)/param(95) > 0 and
 * istep >= that value) it is  r <- dssum(mask r) ; project1 ; hmhzpf (chktcg1 + cggo, 'PRES' -> hmh_gmres) ; project2
 * (navier4.f:513-560,636-1199).  The approximation space {X, B = A X, xbar, bbar, h1old, h2old} is kept ON THE DEVICE, one
 * set per solver name (proj.cuh); `approx` is not touched and napprox(1:2) = (mmx, m) is mirrored back (a caller that sets
 * napprox(2) below the stored m restarts the space).  mxprev = 20 (SIZE.template) => mmx = 8 vectors.
 * nekb_set_projection: INPUT ifprojfld(ifield) and SIZE ldimt_proj (-1 keeps it).  nekb_projection_reset drops all spaces.
 * nekb_hsolve_dev: the same on device pointers (name4: 4 characters, NUL terminated). */
void hsolve_(const char *name, double *u, double *r, const double *h1, const double *h2, const double *vmk, const double *vml,
             const int *imsh, const double *tol, const int *maxit, const int *isd, double *approx, int *napprox, const double *bi,
             size_t name_len);
int nekb_set_projection(int ifield, int ifprojfld, int ldimt_proj);
int nekb_projection_reset(void);
int nekb_hsolve_dev(const char *name4, double *u, double *r, const double *h1, const double *h2, const double *vmk, const double *vml,
                    int imsh, double tol, int maxit, const double *bi, int *napprox, int *niter);

/* ---- Pn-Pn-2 pressure operator E = D (h2 B)^-1 D^T (3-D, lx1 = 8, lx2 = 6) -------------------------
 * core/navier1.f:4095 opgradt(outx,outy,outz,inpfld) -> cdtp (:330-536);  :4064 opdiv(outfld,inpx,inpy,inpz) -> multd
 * (:538-714);  :775 opbinv(out1,out2,out3,inp1,inp2,inp3,h2inv) (inp_i return masked + dssum'ed, as in the reference);
 * :258 cdabdtp(ap,wp,h1,h2,h2inv,intype): intype = 1 -> opbinv, intype = 0 / -1 -> ophinv with (tolhs, nmxv).
 * nekb_set_mesh2 registers what these read from COMMON: ixm12, dxm12 (lx2,lx1 column-major; core/IXYZ, DXYZ), w3m2 (WZ),
 * the nine metric arrays in the order rxm2, sxm2, txm2, rym2, sym2, tym2, rzm2, szm2, tzm2 (GEOM), bm2, bm2inv (MASS; may be
 * NULL when uzawa_gmres is not used), volvm2, tolhs (TSTEP), nmxv (INPUT), nelgv and ifvcor (for ortho). */
int nekb_set_mesh2(int lx2, const double *ixm12, const double *dxm12, const double *w3m2, const double *const *metrics9, const double *bm2,
                   const double *bm2inv, double volvm2, double tolhs, int nmxv, int64_t nelgv, int ifvcor);
void opgradt_(double *outx, double *outy, double *outz, const double *inpfld);
void opdiv_(double *outfld, const double *inpx, const double *inpy, const double *inpz);
void opbinv_(double *out1, double *out2, double *out3, double *inp1, double *inp2, double *inp3, const double *h2inv);
void cdabdtp_(double *ap, const double *wp, const double *h1, const double *h2, const double *h2inv, const int *intype);
/* core/gmres.f:2-237 uzawa_gmres(res,h1,h2,h2inv,intype,iter): right-preconditioned GMRES(lgmres) on E = cdabdtp with
 * hsmg_solve as preconditioner (nekb_hsmg_setup first; param(43) = 0), split weights sqrt(bm2inv) / sqrt(bm2), tolerance
 * through chktcg2 (core/navier1.f:1089-1154).  res is overwritten with the pressure update; iter returns the count.
 * nekb_set_uzawa_state: TSTEP tolps, INPUT param(21), TSTEP prelax, tolpdf. */
void uzawa_gmres_(double *res, const double *h1, const double *h2, const double *h2inv, const int *intype, int *iter);
int nekb_set_uzawa_state(double tolps, double param21, double prelax, double tolpdf);
int nekb_uzawa_gmres_dev(double *res, const double *h1, const double *h2, const double *h2inv, int intype, int *iter, double *div0,
                         double *divex);
int nekb_opgradt_dev(double *ox, double *oy, double *oz, const double *p);
int nekb_opdiv_dev(double *out, const double *ux, const double *uy, const double *uz);
int nekb_cdabdtp_dev(double *ap, const double *wp, const double *h1, const double *h2, const double *h2inv, int intype);

/* INPUT param(idx), 1-based: the path reads param(21) (pressure tolerance), param(22) (Helmholtz tolerance; < 0 relative,
 * core/hmholtz.f:764).  MASS binvm1 / bintm1 (host, lx1^3*nelv / nelt doubles; bintm1 may be NULL) for hmholtz. */
int nekb_set_param(int idx, double value);
int nekb_set_binv(const double *binvm1, const double *bintm1);

/* How the inter-rank part of gs_op runs for this handle: 0 = nothing shared with other ranks, 1 = pack -> grouped
 * ncclSend/ncclRecv -> unpack, 2 = peer-memory exchange (pack kernel stores into the peers' receive areas over NVLink through
 * CUDA IPC mappings and raises epoch flags, unpack kernel waits on them; default on one node, NEKB_GS_P2P=0 selects 1). */
int nekb_gs_exchange_mode(int handle);
/* ------------------------------------------------------------------------------------
 * C. Device-resident API (pointers are device pointers valid on the library's device)
 * ---------------------------------------------------------------------------------- */

int nekb_gs_setup(int *handle, const int64_t *id_host, int64_t n);       /* ids on the host   */
int nekb_gs_setup_dev(int *handle, const int64_t *id_dev, int64_t n);    /* ids on the device */
/* In-place gs_op on a device vector; mask_dev may be NULL, else u is multiplied by the mask after
 * the combine (fuses col2(w,mask), core/hmholtz.f:798 / xmask1 bp5.usr:850). */
int nekb_gs_op_dev(int handle, double *u_dev, int op, const double *mask_dev);
int nekb_gs_free(int handle);
/* Sizes of the handle's local map: groups (ids held by >= 2 local entries) and their members. */
int nekb_gs_info(int handle, int64_t *ngroups, int64_t *nmembers, int64_t *nshared_remote);
/* Copies the handle's CSR (group offsets, member indices, ascending) to host arrays for bit-exact
 * comparison of the gather-scatter index maps. */
int nekb_gs_get_map(int handle, int64_t *off_host, int32_t *idx_host);

/* Remote half of the handle's map (np > 1): number of neighbour ranks and exchange items; then the neighbour
 * ranks (ascending), per-neighbour offsets [npeers+1] and, per item, the local index whose value is sent
 * (items of one neighbour are ordered by ascending global id on both sides). */
int nekb_gs_remote_info(int handle, int *npeers, int64_t *nitems);
int nekb_gs_get_remote(int handle, int *peers, int64_t *peer_off, int32_t *item_rep_idx);

/* ap = A p with registered BP5 geometry; pap_dev (device, may be NULL) receives sum p*ap. */
int nekb_ax_bp5_dev(double *ap_dev, const double *p_dev, double *pap_dev);
int nekb_axhelm_dev(double *au_dev, const double *u_dev, const double *h1_dev, const double *h2_dev,
                    int imesh);
/* dpc = 1/dssum(diag A) on device buffers (core/hmholtz.f:380-524). */
int nekb_setprec_dev(double *dpc_dev, const double *h1_dev, const double *h2_dev, int imesh);
/* Whole cggo solve (core/hmholtz.f:611-846, Jacobi branch) on device buffers.  hist_host (may be NULL,
 * 3*(maxit+2) doubles): per iteration rtz1, rbn2, rho.  *niter = niterhm. */
int nekb_cggo_dev(double *x_dev, const double *f_dev, const double *h1_dev, const double *h2_dev,
                  const double *mask_dev, const double *mult_dev, const double *binv_dev, int imsh, double tin,
                  int maxit, int *niter, double *hist_host);
/* Whole cggos solve on device buffers.  hist_host (may be NULL, 3*maxit doubles) receives per
 * iteration pap, rtz, max|u-x1| (the latter only when tol>0, else 0).  Returns iterations in *niter. */
int nekb_cggos_dev(double *u_dev, const double *rhs_dev, const double *x1_dev, const double *rmult_dev,
                   double tol, int maxit, int *niter, double *hist_host);

/* ------------------------------------------------------------------------------------
 * D. Host-side setup feeding the path
 * ---------------------------------------------------------------------------------- */

/* ---- mesh files: the reference's binary fixtures, read on the host -------------------------
 * .re2: core/reader_re2.f:543-639 header ('#v001' 4-byte words, '#v002'/'#v003' 8-byte words; endian tag 6.54321),
 * :65-158/:391-471 mesh records (group, x(2^ldim), y, z in PREPROCESSOR corner order), :160-290 curved sides,
 * :292-389/:473-541 one boundary-condition section per field (element, face, bl(5), cbl).
 * nekb_re2_info: sizes + the number of records of each BC section (nbc[0..nbc_cap)).
 * nekb_re2_read_mesh: elements [e0,e0+nel) in global order -> xc,yc,zc(2^ldim,nel) (core/INPUT), igroup(nel) (may be NULL).
 * nekb_re2_read_bc: one section -> cbc(3 chars,6,nelgt) and bc(5,6,nelgt), global arrays pre-filled by the caller
 *   (slots without a record are left untouched). */
int nekb_re2_info(const char *path, int64_t *nelgt, int *ldim, int64_t *nelgv, int *wdsize, int64_t *ncurve, int *nsections,
                  int64_t *nbc, int nbc_cap);
int nekb_re2_read_mesh(const char *path, int64_t e0, int64_t nel, double *xc, double *yc, double *zc, int *igroup);
int nekb_re2_read_bc(const char *path, int section, char *cbc, double *bc);
/* nekb_re2_read_curves: the curved-side records -> ccurve(12,nelgt) (one character) and curve(5,12,nelgt) of core/INPUT
 *   (reader_re2.f:160-290 + buf_to_curve :473-497), global arrays pre-filled by the caller (blank / zero). */
int nekb_re2_read_curves(const char *path, char *ccurve, double *curve);
/* .ma2: core/map2.f:712-941 read_map -- 132-byte '#v001' header (7 integers: nel, nactive, depth, d2, npts, nrank,
 * noutflow), endian tag, then 1 + nlv int32 per element: RSB leaf and the vertex ids in SYMMETRIC corner order
 * (-> vertex(nlv,nel) int64 as setupds takes them).  nekb_assign_gllnid: core/map2.f:943-1026 assign_gllnid, in place:
 * leaf -> 0-based rank for np ranks (power-of-two shortcut; otherwise the reference's isort-based contiguous split). */
int nekb_ma2_info(const char *path, int64_t *nel, int64_t *hdr7);
int nekb_ma2_read(const char *path, int nlv, int64_t e0, int64_t nel, int32_t *leaf, int64_t *vertex);
/* .co2: core/map2.f:338-473 read_con -- the connectivity file of a parRSB build: 132-byte '#v001'/'#v002' header (nelgt,
 * nelgv, nv), endian tag, then 1 + nv int32 per element: global element id and vertex ids (-> eid(nel), vertex(nv,nel)
 * int64, the arrays map2.f:206-232 hands to the partitioner and then to setupds).  nlv must equal the file's nv. */
int nekb_co2_info(const char *path, int64_t *nelgt, int64_t *nelgv, int *nv);
int nekb_co2_read(const char *path, int nlv, int64_t e0, int64_t nel, int64_t *eid, int64_t *vertex);
int nekb_assign_gllnid(int *gllnid, int64_t nelgt, int64_t nelgv, int np);

/* core/navier8.f:2004-2360 setvert3d with ifcenter=.false.: glo_num(nx^3,nel) from
 * vertex(8,nel) (symmetric corner order).  Single-process form: np only selects the mod-np
 * bucketing of gbtuple_rank8 (:1964), so the result equals an np-rank reference run.  With a
 * transport registered (nekb_set_transport) and np == nranks the tuples are exchanged for real. */
int nekb_setvert3d(int64_t *glo_num, int64_t *ngv, int nx, int64_t nel, const int64_t *vertex, int np);

/* Host-only: which of this rank's distinct non-zero ids (ascending) also live on other ranks; the rendezvous
 * on rank (id mod np) that gslib's gs_setup performs, over the registered transport.  Call once with
 * peers == NULL to obtain the sizes, then again with arrays: peers[npeers], peer_off[npeers+1],
 * item_ids[nitems] (per neighbour, ascending). */
int nekb_gs_discover(const int64_t *uniq_ids, int64_t n, int *npeers, int64_t *nitems, int *peers,
                     int64_t *peer_off, int64_t *item_ids);

/* Synthetic BP5 case (examples/bp5: genbox.in box, bp5.usr usrdat2/bp5): builds on the device the
 * rank-local part of an nelx*nely*nelz box of [0,1]^3 split into px*py*pz bricks
 * (rank = ix + px*(iy + px*iz); RCB-equivalent of the reference partition),
 * its GLL coordinates, gf, numbering, gs handle, mask, multiplicity, e1 = mask*dsavg(ran1 field),
 * r1 = mask*dssum(A e1).  deform != 0 applies a smooth coordinate perturbation (tests). */
int nekb_bp5_setup(int nelx, int nely, int nelz, int px, int py, int pz, double deform);
/* One cggos solve (bp5.usr:367-369 body) on the device-resident case; returns seconds measured
 * with CUDA events on the library stream (excluded: nothing; the whole call). */
int nekb_bp5_solve(double tol, int maxit, int *niter, double *seconds, double *hist_host);
/* glrdif(u1,e1) of bp5.usr:373,422-447 (global max over ranks when NCCL is up). */
int nekb_bp5_relerr(double *relerr);
/* Host copies of the case's arrays (for parity tests): which = "u1","e1","r1","mask","mult","gf"
 * (reference layout gf(6,lx1^3,nelt)),"xm1","ym1","zm1","bm1","glo_num"(int64).  n_bytes is the buffer capacity. */
int nekb_bp5_get(const char *which, void *host_out, size_t n_bytes);
int64_t nekb_bp5_nel_local(void);
/* Device pointer of one of the case arrays (same names; also "bm1" and "g" = device layout
 * [nelt][6][lx1^3]) for the device-resident API. */
void *nekb_bp5_devptr(const char *which);
int nekb_bp5_gs_handle(void);

/* ------------------------------------------------------------------------------------
 * F. Pressure preconditioner (additive Schwarz / FDM multigrid) and its GMRES driver
 * ---------------------------------------------------------------------------------- */

/* core/hsmg.f:2234 h1mg_setup() (+ core/navier6.f:82 swap_lengths, core/navier8.f:83 set_up_h1_crs).  The
 * reference routine has no arguments and reads COMMON state; the Fortran glue passes that state here:
 *   fbc[6*nelv]  (lbr,rbr,lbs,rbs,lbt,rbt) of get_fast_bc (core/fast3d.f:802-877) per element: 0 interior /
 *                periodic, 1 Dirichlet for the pressure ('O','ON',...), 2 Neumann ('v','W','SYM',...);
 *   xm1,ym1,zm1  /gxyz/ coordinates (host, lx1^3*nelv each) for swap_lengths / plane_space;
 *   vertex       /ivrtx/ vertex(8,nelv) in symmetric order (get_vert);
 *   null_space   ifvcor (core/navier8.f:208-212).
 * The geometric factors must have been registered (section B); levels follow h1mg_setup_mg_nx (:2272-2337). */
int nekb_h1mg_setup(const int *fbc, const double *xm1, const double *ym1, const double *zm1, const int64_t *vertex,
                    int nelv, int null_space);
/* core/hsmg.f:1855 h1mg_solve(z,rhs,if_hybrid), additive form (if_hybrid = .false., core/gmres.f:330), on device
 * buffers.  As in the reference, rhs is masked in place (hsmg.f:449). */
int nekb_h1mg_solve_dev(double *z_dev, double *rhs_dev);
void h1mg_solve_(double *z, double *rhs, const int *if_hybrid);
/* One h1mg_schwarz (core/hsmg.f:425-494, sigma = 1) at `level` (1-based as in the reference, 2..lmax); r is masked in
 * place.  And the coarse solve of hsmg_coarse_solve (:1321-1354 -> crs_solve, core/crs_xxt.c:926-965) on the
 * 2^3-per-element vertex arrays. */
int nekb_h1mg_schwarz_dev(int level, double *e_dev, double *r_dev);
int nekb_crs_solve_dev(double *e_dev, const double *r_dev);
/* Sizes: number of levels, points per direction of every level (coarse first), distinct 1-D eigen-systems per
 * level, iterations of the last coarse solve. */
int nekb_h1mg_info(int *lmax, int *nh3, int *ntab3, int *crs_iters);
/* The reference's coarse-solver facade, core/fcrs.c:45-96 (crs_setup / crs_solve / crs_free over core/crs_xxt.c:860-965), as
 * called from core/navier8.f:217 (set_up_h1_crs) and core/hsmg.f:1349, navier8.f:53,1528: n local dofs with global ids id[]
 * (0 = ignored: Dirichlet, set_jl_crs_mask), the local operator as nz COO entries over 0-based local dofs (set_mat_ij),
 * null_space flag.  Only sid = 0 (XXT, param(40) = 0) is provided -- a direct solve: the assembled matrix of all ranks is
 * inverted on the device once, a solve is gather + GEMV + scatter; with a null space the result is mean-free over the
 * distinct dofs like crs_xxt.c:951-960.  comm / np / param / datafname are ignored (the library's own transport).
 * Limited to NEKB_CRS_DENSE_MAX (12288) distinct dofs.  Status: compiled, parity test written, NOT YET RUN ON A GPU. */
void crs_setup_(int *handle, const int *sid, const int *comm, const int *np, const int *n, const int64_t *id, const int *nz,
                const int *Ai, const int *Aj, const double *A, const int *null_space, const double *param,
                const char *datafname, int *ierr);
void crs_solve_(const int *handle, double *x, const double *b);
void crs_free_(const int *handle);
int nekb_fcrs_solve_dev(int handle, double *x_dev, const double *b_dev);
/* HOST set-up only (nek5000_b200/csrc/crs_amg.cuh; no device work, nothing in the solve path uses it yet): an aggregation
 * hierarchy for coarse problems beyond the dense limit -- the role of crs_xxt.c / crs_amg.c's set-up.  Input: the assembled
 * operator as COO triplets over 0-based dofs (duplicates are summed in input order); levels are built until <= nmax rows
 * (greedy aggregation with strength threshold theta, piecewise-constant prolongation, Galerkin products).  level_get copies
 * a level's CSR (rowptr[n+1], col[nnz], val[nnz]) and, for every level but the coarsest, its aggregate map agg[n]. */
int nekb_crs_amg_build_host(int64_t n, int64_t nz, const int64_t *I, const int64_t *J, const double *V, int64_t nmax, double theta,
                            double omega_p, int *nlevels);
/* omega_p = 0: piecewise-constant prolongation; > 0: smoothed aggregation, P = (I - omega_p D^-1 A) P_tentative.
 * level_p: the prolongation of a level (n_level rows, n_{level+1} columns) in CSR; pass NULL arrays to query nnz first. */
int nekb_crs_amg_level_p(int level, int64_t *nnz, int64_t *rowptr, int32_t *col, double *val);
int nekb_crs_amg_level_info(int level, int64_t *n, int64_t *nnz);
int nekb_crs_amg_level_get(int level, int64_t *rowptr, int32_t *col, double *val, int32_t *agg);
/* Device side of the same (csrc/crs_amg_dev.cuh; compiled, NOT YET RUN ON A GPU, not used by h1mg_solve): upload moves the
 * host hierarchy to the device (damped-Jacobi weight omega) and inverts the coarsest operator; solve_dev runs CG on the
 * finest operator, preconditioned by one V(1,1) cycle, to a relative 2-norm residual tol (SPD systems). */
int nekb_crs_amg_upload(double omega);
int nekb_crs_amg_solve_dev(double *x_dev, const double *b_dev, double tol, int maxit, int *iters);
/* Host copies of setup products for parity tests.  which: "mask","rstr_wt","swt" (level-sized, level 1-based),
 * "J" (interpolation level -> level+1, row-major nf x nc), "lm","ll","lr" (3*nelv, direction-major; level ignored),
 * "crs_a" (64*nelv, a(i,j,e) as a[e][i][j]). */
int nekb_h1mg_get(const char *which, int level, double *host_out, size_t n_doubles);
/* Relative residual and iteration cap of the device coarse solve (defaults 1e-13, 2000). */
int nekb_crs_set_tolerance(double tol, int maxit);
void nekb_h1mg_free(void);

/* Pn-Pn-2 pressure preconditioner: core/hsmg.f:22-47 hsmg_setup, :1376-1602 hsmg_solve(e,r) and its top level
 * core/fasts.f:2-94 local_solves_fdm(u,v), on the lx2 = lx1-2 Gauss grid (arrays of (lx1-2)^3*nelv doubles).
 * The levels below the top, the coarse solve and all exchanges are built here exactly as for h1mg.  The top-level
 * fast-diagonalisation data (common /fastd/: df(lx1^3,nelv), sr/ss/st(2*lx1^2,nelv), S in the first half of each,
 * column-major; host arrays) is either REGISTERED -- what gen_fast (core/fast3d.f:2-140) left in COMMON -- or, when all
 * four pointers are NULL, COMPUTED here: gen_fast with param(44) = 0 (set_up_fast_1D_sem :1351-1408, _sem_op :1410-1540,
 * load_semhat_weighted :1181-1213) from the lengths of swap_lengths and the fbc codes (0 element, 1 outflow, 2 wall,
 * 3 symmetry).
 * nelgv: global element count (ortho, core/navier1.f:223).  fbc, xm1.., vertex as for nekb_h1mg_setup. */
int nekb_hsmg_setup(const int *fbc, const double *xm1, const double *ym1, const double *zm1, const int64_t *vertex,
                    int nelv, int null_space, int64_t nelgv, const double *df, const double *sr, const double *ss,
                    const double *st);
/* Host-only (no GPU needed): the 1-D generalised eigen-systems the preconditioner setup is built from, for the CPU tests.
 * nekb_fast1d_sem_host: gen_fast's set_up_fast_1D_sem (core/fast3d.f:1351-1408) at order lx1-1 -> S[lx1*lx1] (row-major,
 * eigenvectors in columns, boundary rows zeroed), lam[lx1].  nekb_fast1d_host: hsmg_setup_fast1d (core/hsmg.f:775-879) for
 * polynomial order n -> S[(n+3)^2], lam[n+3].  bc codes of get_fast_bc. */
int nekb_fast1d_sem_host(int lx1, int lbc, int rbc, double ll, double lm, double lr, double *S, double *lam);
int nekb_fast1d_host(int n, int lbc, int rbc, double ll, double lm, double lr, double *S, double *lam);
int nekb_hsmg_solve_dev(double *e_dev, const double *r_dev);
int nekb_local_solves_fdm_dev(double *u_dev, const double *v_dev);
void hsmg_solve_(double *e, const double *r);
void local_solves_fdm_(double *u, const double *v);
/* Host copies for parity tests: "J" (level -> level+1), "owt" (top-level overlap weight), "swt", "mask". */
int nekb_hsmg_get(const char *which, int level, double *host_out, size_t n_doubles);

/* Single-level Schwarz / FDM preconditioner on the lx1^3 tiles (core/FDMH1).
 * nekb_fdm_h1_setup = set_fdm_prec_h1A (core/hmholtz.f:1028-1220) for one field: face_internal[6*nel] is 1 where cbc is
 * 'E  ','P  ','p  ' (faces r-,r+,s-,s+,t-,t+), mask the field's Dirichlet mask, xm1..zm1 /gxyz/ (all host arrays).
 * nekb_set_kfldfdm = common /fdmh1i/ kfldfdm: >= 0 makes cggo take its Schwarz branch (:686-691, :731-746). */
int nekb_fdm_h1_setup(const int *face_internal, const double *mask, const double *xm1, const double *ym1, const double *zm1,
                      int nel);
int nekb_set_kfldfdm(int kfldfdm);
/* core/hmholtz.f:1222 set_fdm_prec_h1b(d,h1,h2,nel) and :937 fdm_h1(z,r,d,mask,mult,nel,kt,rr).  kt is not read: it is
 * ktype(1,1,kfldfdm) of the same COMMON the setup call registered; mult and rr are unused by the reference too. */
int nekb_set_fdm_prec_h1b_dev(double *d_dev, const double *h1_dev, const double *h2_dev);
int nekb_fdm_h1_dev(double *z_dev, const double *r_dev, const double *d_dev, const double *mask_dev);
void set_fdm_prec_h1b_(double *d, const double *h1, const double *h2, const int *nel);
void fdm_h1_(double *z, const double *r, const double *d, const double *mask, const double *mult, const int *nel, const int *kt,
             double *rr);
/* Host copies of the setup products: which = "ktype" (int32 [nel][3], 1..9 as in the reference), "elsize" (double
 * [nel][3]), "dd" (double [9][lx1]). */
int nekb_fdm_h1_get(const char *which, void *host_out, size_t n_bytes);

/* State hmh_gmres reads from COMMON: pmask (core/SOLN), binvm1 (core/MASS), tolps (core/TSTEP), param(21), ifvcor
 * (core/INPUT), nelgv; volvm1 and istep come from nekb_set_step_info.  Host arrays of lx1^3*nelv doubles. */
int nekb_set_pressure_state(const double *pmask, const double *binvm1, double tolps, double param21, int ifvcor,
                            int64_t nelgv);
/* core/gmres.f:304 hmh_gmres(res,h1,h2,wt,iter): on entry *iter = maxit, on return the iterations performed; res is
 * overwritten with the solution.  h1mg_solve is the preconditioner (ifmgrid, param(40) in 0..2). */
void hmh_gmres_(double *res, const double *h1, const double *h2, const double *wt, int *iter);
/* core/hmholtz.f:2164 hmh_flex_cg(res,h1,h2,wt,iter): flexible PCG with h1mg_solve as preconditioner (param(42) = 2); the same
 * COMMON state as hmh_gmres.  cggo_/hmholtz_/hsolve_ forward name = 'PRES' here when param(42) = 2. */
void hmh_flex_cg_(double *res, const double *h1, const double *h2, const double *wt, int *iter);
/* The same on device buffers with an explicit tolerance: tol > 0 absolute on |gamma|/sqrt(volvm1), tol < 0 relative
 * to the initial residual (param(21) < 0).  hist_host (may be NULL, maxit+1 doubles) receives rnorm per iteration;
 * h2_dev may be NULL (h2 = 0).  ifvcor/nelgv as registered by nekb_set_pressure_state. */
int nekb_hmh_gmres_dev(double *res_dev, const double *h1_dev, const double *h2_dev, const double *wt_dev,
                       const double *pmask_dev, double tol, int maxit, int *iter, double *hist_host, double *div0);

/* ------------------------------------------------------------------------------------
 * E. Plain device-memory helpers for host programs without a CUDA binding of their own
 * ---------------------------------------------------------------------------------- */
void *nekb_dev_alloc(size_t bytes);
void nekb_dev_free(void *dev);
int nekb_h2d(void *dev, const void *host, size_t bytes);
int nekb_d2h(void *host, const void *dev, size_t bytes);
int nekb_d2d(void *dst_dev, const void *src_dev, size_t bytes);
int nekb_sync(void);

#ifdef __cplusplus
}
#endif
#endif /* NEKB200_H */
