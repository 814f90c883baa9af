// nekb200.cu -- C-ABI of libnekb200.so (see include/nekb200.h).  Unity build: the kernels live in the
// headers included below.  Build: nvcc -gencode arch=compute_100a,code=sm_100a (nek5000_b200/build.py).
#include "../../include/nekb200.h"

#include "gmres.cuh"
#include "hcg.cuh"
#include "proj.cuh"
#include "pnpn2.cuh"
#include "readers.cuh"
#include "crs_amg_dev.cuh"

using namespace nekb;

namespace {

template <class F>
int guard(F &&f)
{
    try {
        f();
        return 0;
    } catch (const std::exception &e) {
        ctx().last_error = e.what();
        return 1;
    }
}

// Fortran-named entry points have no status argument: report and leave through the exit handler
// (reference: exitt, core/comm_mpi.f:550-636); never return with stale output.
template <class F>
void guard_fortran(const char *who, F &&f)
{
    if (guard(f)) {
        fprintf(stderr, "nekb200: %s: %s\n", who, ctx().last_error.c_str());
        fflush(stderr);
        if (ctx().exit_handler)
            ctx().exit_handler();
        else
            abort();
    }
}

inline int64_t field_len() { return (int64_t)ctx().nelt * ctx().nxyz; }

// interleaved gf(6,nxyz,nel) -> g[e][c][q]
__global__ void __launch_bounds__(256) gf_deinterleave_kernel(double *__restrict__ g, const double *__restrict__ gf, int nxyz, int64_t total)
{
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
        const int64_t e = t / (6 * nxyz);
        const int r = (int)(t - e * 6 * nxyz), cidx = r / nxyz, q = r % nxyz;
        g[t] = gf[(e * nxyz + q) * 6 + cidx];
    }
}
__global__ void __launch_bounds__(256) gf_interleave_kernel(double *__restrict__ gf, const double *__restrict__ g, int nxyz, int64_t total)
{
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
        const int64_t e = t / (6 * nxyz);
        const int r = (int)(t - e * 6 * nxyz), q = r / 6, cidx = r % 6;
        gf[t] = g[(e * 6 + cidx) * nxyz + q];
    }
}
// elements that are not deformed use g1..g3 only (core/hmholtz.f:196-206): clear their cross terms
__global__ void zero_cross_terms_kernel(double *__restrict__ g, const int *__restrict__ dfrm, int nxyz, int nel)
{
    const int e = blockIdx.x;
    if (e >= nel || dfrm[e]) return;
    double *ge = g + (size_t)e * 6 * nxyz;
    for (int q = threadIdx.x; q < nxyz; q += blockDim.x) ge[1 * nxyz + q] = 0.0, ge[2 * nxyz + q] = 0.0, ge[4 * nxyz + q] = 0.0;
}

__global__ void __launch_bounds__(CG_THREADS)
    glsc3_kernel(const double *__restrict__ a, const double *__restrict__ b, const double *__restrict__ m, int64_t n,
                 double *out, double *partials, unsigned *counter)
{
    __shared__ double red[33];
    double s = 0.0;
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < n; t += (int64_t)gridDim.x * blockDim.x)
        s = fma(a[t] * b[t], m[t], s);
    double bs = block_reduce(s, red);
    grid_reduce(bs, partials, counter, red, [=](double tot) { *out = tot; });
}

__global__ void __launch_bounds__(CG_THREADS)
    glrdif_kernel(const double *__restrict__ x, const double *__restrict__ y, int64_t n, double *out3, double *partials,
                  unsigned *counter)
{
    __shared__ double red[33];
    double d = 0.0, xm = 0.0, ym = 0.0;
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < n; t += (int64_t)gridDim.x * blockDim.x) {
        d = fmax(d, fabs(x[t] - y[t]));
        xm = fmax(xm, x[t]);
        ym = fmax(ym, y[t]);
    }
    double b0 = block_reduce<true>(d, red), b1 = block_reduce<true>(xm, red), b2 = block_reduce<true>(ym, red);
    grid_reduce<true>(b0, partials, counter, red, [=](double t) { out3[0] = t; });
    grid_reduce<true>(b1, partials + CG_PART_STRIDE, counter + 1, red, [=](double t) { out3[1] = t; });
    grid_reduce<true>(b2, partials + 2 * CG_PART_STRIDE, counter + 2, red, [=](double t) { out3[2] = t; });
}

void apply_ifdfrm()
{
    ctx().geom_gen++;   // every registration of factors ends here
    Ctx &c = ctx();
    if (!c.have_geom || c.ifdfrm.empty()) return;
    NEKB_REQUIRE((int)c.ifdfrm.size() >= c.nelt, "ifdfrm shorter than nelt");
    DevBuf<int> f;
    f.upload(c.ifdfrm.data(), (size_t)c.nelt, c.stream);
    zero_cross_terms_kernel<<<c.nelt, 128, 0, c.stream>>>(c.g.p, f.p, c.nxyz, c.nelt);
    NEKB_LAUNCHED();
    NEKB_CUDA(cudaStreamSynchronize(c.stream));
}

int field_handle()
{
    Ctx &c = ctx();
    NEKB_REQUIRE(c.ifield >= 0 && c.ifield < 32, "ifield out of range");
    const int h = c.gsh_fld[c.ifield];
    NEKB_REQUIRE(h >= 0, "no gs handle registered for the current ifield (nekb_set_field_handle / setupds_)");
    return h;
}

}  // namespace

namespace nekb {
int gs_setup_from_host_ids(const int64_t *id_host, int64_t n, const int32_t *cand, int64_t ncand)
{
    Ctx &c = ctx();
    const int hnd = gs_new_handle();
    GsMap &h = c.gs[hnd];
    DevBuf<int64_t> ids;
    ids.upload(id_host, (size_t)n, c.stream);
    gs_build_local(h, ids.p, n);
    gs_build_remote(h, id_host, n, cand, ncand);
    return hnd;
}
}  // namespace nekb

extern "C" {

// ---------------------------------------------------------------------------------------------------- lifecycle
int nekb_init(int device, int lx1, int ldim)
{
    return guard([&] {
        Ctx &c = ctx();
        NEKB_REQUIRE(ldim == 3, "only ldim = 3 is supported");
        NEKB_REQUIRE(lx1 >= 2 && lx1 <= MAX_NX, "lx1 out of range");
        if (c.inited) {
            NEKB_REQUIRE(c.device == device && c.nx == lx1, "nekb_init called again with different arguments");
            return;
        }
        NEKB_CUDA(cudaSetDevice(device));
        cudaDeviceProp prop;
        NEKB_CUDA(cudaGetDeviceProperties(&prop, device));
        c.device = device;
        c.num_sms = prop.multiProcessorCount;
        c.nx = lx1;
        c.nxyz = lx1 * lx1 * lx1;
        NEKB_CUDA(cudaStreamCreateWithFlags(&c.stream, cudaStreamNonBlocking));
        NEKB_CUDA(cudaEventCreate(&c.ev0));
        NEKB_CUDA(cudaEventCreate(&c.ev1));
        c.sc.alloc(1);
        c.sc.zero(c.stream);
        c.partials.alloc(4 * CG_PART_STRIDE);
        for (int i = 0; i < 32; i++) c.gsh_fld[i] = -1;
        NEKB_CUDA(cudaStreamSynchronize(c.stream));
        c.inited = true;
    });
}

void nekb_finalize(void)
{
    Ctx &c = ctx();
    if (!c.inited) return;
    try {
        comm_quiesce();   // peers may still be raising flags in this rank's exchange memory
    } catch (...) {
    }
    bp5case() = Bp5Case();
    crs_release_graph();
    h1mg() = H1mg();
    hsmg2() = H1mg();
    fcrs_table().clear();
    amg_dev() = AmgDev();
    fdm_h1_state() = FdmH1State();
    gmres_state() = GmresState();
    crs_scalars().release();
    hcg_state() = HcgState();
    proj_states().clear();
    proj_scratch() = ProjScratch();
    mesh2() = Mesh2();
    pnpn2_work() = Pnpn2Work();
    c.gs.clear();
    if (c.nccl_comm) nccl().CommDestroy(comm_handle());
    void (*eh)(void) = c.exit_handler;
    cudaStream_t s = c.stream;
    cudaEvent_t e0 = c.ev0, e1 = c.ev1;
    c = Ctx();
    c.exit_handler = eh;
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaStreamDestroy(s);
}

const char *nekb_last_error(void) { return ctx().last_error.c_str(); }
void nekb_set_exit_handler(void (*handler)(void)) { ctx().exit_handler = handler; }
void *nekb_stream(void) { return (void *)ctx().stream; }
int64_t nekb_launch_count(int reset)
{
    const int64_t v = launch_counter();
    if (reset) launch_counter() = 0;
    return v;
}

int nekb_prof_enable(int on)
{
    Prof &p = prof();
    p.on = on != 0;
    for (int k = 0; k < PROF_NCAT; k++) p.secs[k] = 0.0, p.count[k] = 0;
    p.spans.clear();
    p.used = 0;
    return 0;
}
int nekb_prof_get(const char *kernel, double *seconds, int64_t *launches)
{
    return guard([&] {
        const std::string w(kernel);
        int cat = -1;
        if (w == "ax") cat = PROF_AX;
        if (w == "gs") cat = PROF_GS;
        if (w == "update") cat = PROF_UPDATE;
        if (w == "pupdate") cat = PROF_PUPDATE;
        NEKB_REQUIRE(cat >= 0, "unknown kernel class '" + w + "' (ax, gs, update, pupdate)");
        if (seconds) *seconds = prof().secs[cat];
        if (launches) *launches = prof().count[cat];
    });
}

// ---------------------------------------------------------------------------------------------------- multi-rank
int nekb_set_transport(int rank, int nranks, nekb_allgather_fn allgather, nekb_alltoallv_fn alltoallv, void *user)
{
    return guard([&] {
        NEKB_REQUIRE(nranks >= 1 && rank >= 0 && rank < nranks, "bad rank / nranks");
        Ctx &c = ctx();
        c.rank = rank, c.nranks = nranks;
        c.allgather = allgather, c.alltoallv = alltoallv, c.transport_user = user;
    });
}
int nekb_comm_unique_id(void *id_out_128)
{
    return guard([&] {
        nccl_unique_id id;
        NEKB_NCCL(nccl().GetUniqueId(&id));
        memcpy(id_out_128, &id, sizeof id);
    });
}
int nekb_comm_init(const void *id_in_128, int rank, int nranks)
{
    return guard([&] {
        require_init();
        Ctx &c = ctx();
        NEKB_REQUIRE(nranks >= 1 && rank >= 0 && rank < nranks, "bad rank / nranks");
        nccl_unique_id id;
        memcpy(&id, id_in_128, sizeof id);
        nccl_comm_t comm = nullptr;
        NEKB_NCCL(nccl().CommInitRank(&comm, nranks, id, rank));
        c.nccl_comm = comm;
        c.rank = rank, c.nranks = nranks;
    });
}

// ---------------------------------------------------------------------------------------------------- registration
int nekb_set_nel(int nelv, int nelt)
{
    return guard([&] {
        NEKB_REQUIRE(nelv >= 0 && nelt >= nelv, "need 0 <= nelv <= nelt");
        ctx().nelv = nelv, ctx().nelt = nelt;
    });
}
int nekb_set_gll(const double *zgm1, const double *wxm1)
{
    return guard([&] {
        require_init();
        Ctx &c = ctx();
        c.z_host.assign(zgm1, zgm1 + c.nx);
        c.w_host.assign(wxm1, wxm1 + c.nx);
        c.have_gll = true;
    });
}
int nekb_set_dxyz(const double *dxm1, const double *dxtm1)
{
    return guard([&] {
        require_init();
        Ctx &c = ctx();
        const int nx = c.nx;
        c.D_host.resize((size_t)nx * nx);
        for (int a = 0; a < nx; a++)
            for (int b = 0; b < nx; b++) {
                c.D_host[(size_t)a * nx + b] = dxm1[a + nx * b];  // Fortran dxm1(a,b)
                NEKB_REQUIRE(dxtm1 == nullptr || dxtm1[b + nx * a] == dxm1[a + nx * b], "dxtm1 is not the transpose of dxm1");
            }
        NEKB_CUDA(cudaMemcpyToSymbolAsync(c_D, c.D_host.data(), sizeof(double) * nx * nx, 0, cudaMemcpyHostToDevice, c.stream));
        NEKB_CUDA(cudaStreamSynchronize(c.stream));
        c.have_D = true;
    });
}
int nekb_set_geom(const double *g1m1, const double *g2m1, const double *g3m1, const double *g4m1, const double *g5m1,
                  const double *g6m1, const double *bm1)
{
    return guard([&] {
        require_init();
        Ctx &c = ctx();
        NEKB_REQUIRE(c.nelt > 0, "call nekb_set_nel first");
        const size_t nxyz = c.nxyz, nel = c.nelt;
        c.g.alloc(6 * nxyz * nel);
        // core order g1..g6 = rr,ss,tt,rs,rt,st -> device slots rr,rs,rt,ss,st,tt
        const double *src[6] = {g1m1, g4m1, g5m1, g2m1, g6m1, g3m1};
        for (int k = 0; k < 6; k++)
            NEKB_CUDA(cudaMemcpy2DAsync(c.g.p + k * nxyz, 6 * nxyz * sizeof(double), src[k], nxyz * sizeof(double),
                                        nxyz * sizeof(double), nel, cudaMemcpyHostToDevice, c.stream));
        c.bm1.upload(bm1, nxyz * nel, c.stream);
        NEKB_CUDA(cudaStreamSynchronize(c.stream));
        c.have_geom = true;
        apply_ifdfrm();
    });
}
int nekb_set_geom_bp5(const double *gf)
{
    return guard([&] {
        require_init();
        Ctx &c = ctx();
        NEKB_REQUIRE(c.nelt > 0, "call nekb_set_nel first");
        const int64_t total = (int64_t)6 * c.nxyz * c.nelt;
        DevBuf<double> tmp;
        tmp.upload(gf, (size_t)total, c.stream);
        c.g.alloc((size_t)total);
        gf_deinterleave_kernel<<<blocks_for(total), 256, 0, c.stream>>>(c.g.p, tmp.p, c.nxyz, total);
        NEKB_LAUNCHED();
        NEKB_CUDA(cudaStreamSynchronize(c.stream));
        c.have_geom = true;
        c.geom_gen++;
    });
}
int nekb_set_geom_from_xyz(const double *xm1, const double *ym1, const double *zm1, int bp5_form)
{
    return guard([&] {
        require_init();
        Ctx &c = ctx();
        NEKB_REQUIRE(c.nelt > 0, "call nekb_set_nel first");
        const size_t n = (size_t)field_len();
        DevBuf<double> x, y, z;
        x.upload(xm1, n, c.stream), y.upload(ym1, n, c.stream), z.upload(zm1, n, c.stream);
        const int nelv = c.nelv;
        geom_from_xyz(x.p, y.p, z.p, c.nelt, bp5_form ? 0 : 1, nullptr);
        c.nelv = nelv;
        NEKB_CUDA(cudaStreamSynchronize(c.stream));
        apply_ifdfrm();
    });
}
int nekb_get_geom(double *g1m1, double *g2m1, double *g3m1, double *g4m1, double *g5m1, double *g6m1, double *bm1,
                  double *gf)
{
    return guard([&] {
        require_init();
        Ctx &c = ctx();
        NEKB_REQUIRE(c.have_geom, "geometry not registered");
        const size_t nxyz = c.nxyz, nel = c.nelt;
        double *dst[6] = {g1m1, g4m1, g5m1, g2m1, g6m1, g3m1};
        for (int k = 0; k < 6; k++)
            if (dst[k])
                NEKB_CUDA(cudaMemcpy2DAsync(dst[k], nxyz * sizeof(double), c.g.p + k * nxyz, 6 * nxyz * sizeof(double),
                                            nxyz * sizeof(double), nel, cudaMemcpyDeviceToHost, c.stream));
        if (bm1 && c.bm1.n >= nxyz * nel) NEKB_CUDA(cudaMemcpyAsync(bm1, c.bm1.p, nxyz * nel * sizeof(double), cudaMemcpyDeviceToHost, c.stream));
        if (gf) {
            const int64_t total = (int64_t)6 * nxyz * nel;
            DevBuf<double> tmp;
            tmp.alloc((size_t)total);
            gf_interleave_kernel<<<blocks_for(total), 256, 0, c.stream>>>(tmp.p, c.g.p, c.nxyz, total);
            NEKB_LAUNCHED();
            tmp.download(gf, (size_t)total, c.stream);
        }
        NEKB_CUDA(cudaStreamSynchronize(c.stream));
    });
}
int nekb_set_ifdfrm(const int *ifdfrm)
{
    return guard([&] {
        Ctx &c = ctx();
        if (ifdfrm == nullptr) {
            c.ifdfrm.clear();
            return;
        }
        NEKB_REQUIRE(c.nelt > 0, "call nekb_set_nel first");
        c.ifdfrm.assign(ifdfrm, ifdfrm + c.nelt);
        apply_ifdfrm();
    });
}
int nekb_set_v1mask(const double *v1mask)
{
    return guard([&] {
        require_init();
        ctx().v1mask.upload(v1mask, (size_t)field_len(), ctx().stream);
        NEKB_CUDA(cudaStreamSynchronize(ctx().stream));
    });
}
int nekb_set_ifield(int ifield)
{
    return guard([&] {
        NEKB_REQUIRE(ifield >= 0 && ifield < 32, "ifield out of range");
        ctx().ifield = ifield;
    });
}
int nekb_ax_affine_active(void)
{
    int v = 0;
    guard([&] { v = ax_affine_ensure() ? 1 : 0; });
    return v;
}
double nekb_ax_affine_deviation(void) { return ctx().affine_maxdev; }
int nekb_last_history(double *out, int64_t capacity, int *rows, int *cols)
{
    return guard([&] {
        Ctx &c = ctx();
        if (rows) *rows = c.last_hist_rows;
        if (cols) *cols = c.last_hist_cols;
        const int64_t want = (int64_t)c.last_hist_rows * c.last_hist_cols;
        if (out) {
            NEKB_REQUIRE(capacity >= want, "nekb_last_history: output buffer too small");
            memcpy(out, c.last_hist.data(), sizeof(double) * (size_t)want);
        }
    });
}
int nekb_set_restol(int ifield, double restol)
{
    return guard([&] {
        NEKB_REQUIRE(ifield >= 0 && ifield < 32, "ifield out of range");
        NEKB_REQUIRE(restol >= 0.0, "restol must be >= 0 (0 = cggo uses the caller's tolerance)");
        ctx().restol[ifield] = restol;
    });
}
int nekb_set_field_handle(int ifield, int gs_handle)
{
    return guard([&] {
        NEKB_REQUIRE(ifield >= 0 && ifield < 32, "ifield out of range");
        gs_get(gs_handle);
        ctx().gsh_fld[ifield] = gs_handle;
    });
}
int nekb_set_step_info(int istep, double volvm1, double voltm1)
{
    ctx().istep = istep, ctx().volvm1 = volvm1, ctx().voltm1 = voltm1;
    return 0;
}
int nekb_set_param(int idx, double value)
{
    return guard([&] {
        NEKB_REQUIRE(idx >= 1 && idx <= 200, "nekb_set_param: index out of range (1..200)");
        ctx().param[idx] = value;
    });
}
int nekb_set_binv(const double *binvm1, const double *bintm1)
{
    return guard([&] {
        require_init();
        Ctx &c = ctx();
        if (binvm1) c.binvm1.upload(binvm1, (size_t)c.nelv * c.nxyz, c.stream);
        if (bintm1) c.bintm1.upload(bintm1, (size_t)c.nelt * c.nxyz, c.stream);
        NEKB_CUDA(cudaStreamSynchronize(c.stream));
    });
}
int nekb_niterhm(void) { return ctx().niterhm; }
int nekb_set_velocity_state(const double *v1mask, const double *v2mask, const double *v3mask, const double *vmult)
{
    return guard([&] {
        require_init();
        Ctx &c = ctx();
        const size_t n = (size_t)c.nelv * c.nxyz;
        const double *m[3] = {v1mask, v2mask, v3mask};
        for (int k = 0; k < 3; k++)
            if (m[k]) c.vmask[k].upload(m[k], n, c.stream);
        if (vmult) c.vmult.upload(vmult, n, c.stream);
        NEKB_CUDA(cudaStreamSynchronize(c.stream));
    });
}
int nekb_niterhm3(int *niter3)
{
    return guard([&] {
        for (int k = 0; k < 3; k++) niter3[k] = ctx().niter3[k];
    });
}

// Three Helmholtz solves sharing h1, h2 (ophinv's body): rhs_c <- mask_c * dssum(rhs_c) in place, tolerance per component
// through chktcg1 exactly as hmholtz does, then the fused 3-right-hand-side PCG (hcg.cuh) or, when a branch it does not
// provide is needed, cggo_run component by component.  All pointers are device pointers.
static void ophinv_dev(double *const *o, double *const *rhs, const double *h1, const double *h2, const double *const *mask,
                       const double *mult, const double *binv, double tolh, int maxit, int *niter3, double *hist_host)
{
    Ctx &c = ctx();
    const int nel = c.nelv;
    const int64_t n = (int64_t)nel * c.nxyz;
    NEKB_REQUIRE(c.volvm1 > 0.0, "volvm1 not registered (nekb_set_step_info)");
    // ifh2 decides whether chktcg1's axhelm carries the mass term
    absmax_kernel<<<cg_grid(n), CG_THREADS, 0, c.stream>>>(h2, n, &c.sc.p->work[3], c.partials.p, &c.sc.p->counter[0]);
    NEKB_LAUNCHED();
    comm_allreduce_max(&c.sc.p->work[3], 1);
    double h2max = 0.0;
    NEKB_CUDA(cudaMemcpyAsync(&h2max, &c.sc.p->work[3], sizeof(double), cudaMemcpyDeviceToHost, c.stream));
    NEKB_CUDA(cudaStreamSynchronize(c.stream));
    double tol[3];
    for (int k = 0; k < 3; k++) {
        gs_op(field_handle(), rhs[k], 1, mask[k]);                                   // hmholtz.f:55-56
        tol[k] = fabs(tolh);
        if (c.param[22] == 0.0 || c.istep <= 10)                                     // :59-60
            tol[k] = chktcg1_dev(tol[k], rhs[k], h1, h2max > 0.0 ? h2 : nullptr, mask[k], mult, binv, nel, c.volvm1);
        if (tolh < 0) tol[k] = tolh;                                                 // :62
        tol[k] = cggo_tin(tol[k]);                                                   // :676 restol(ifield)
    }
    const double *f[3] = {rhs[0], rhs[1], rhs[2]};
    if (hcg_applicable(3) &&
        hcg_run(3, o, f, h1, h2, mask, mult, binv, field_handle(), nel, c.volvm1, tol, maxit, c.istep, niter3, hist_host))
        return;
    for (int k = 0; k < 3; k++) {
        CggoArgs a{o[k], rhs[k], h1, h2, mask[k], mult, binv, field_handle(), nel, c.volvm1, c.istep};
        niter3[k] = cggo_run(a, tol[k], maxit, nullptr);
    }
}
int nekb_ophinv_dev(double *o1, double *o2, double *o3, double *i1, double *i2, double *i3, const double *h1, const double *h2,
                    const double *m1, const double *m2, const double *m3, const double *mult, const double *binv, double tolh, int maxit,
                    int *niter3, double *hist_host)
{
    return guard([&] {
        require_init();
        double *o[3] = {o1, o2, o3}, *r[3] = {i1, i2, i3};
        const double *m[3] = {m1, m2, m3};
        int it[3] = {0, 0, 0};
        ophinv_dev(o, r, h1, h2, m, mult, binv, tolh, maxit, it, hist_host);
        for (int k = 0; k < 3; k++) ctx().niter3[k] = it[k];
        ctx().niterhm = it[2];
        if (niter3)
            for (int k = 0; k < 3; k++) niter3[k] = it[k];
    });
}
// core/induct.f:1022-1090 ophinv(o1,o2,o3,i1,i2,i3,h1,h2,tolh,nmxhi): the standard branch (ifstrs = .false., no residual
// projection: param(93) = 0 or ifprojfld(ifield) = .false.), ifield = 1.  v1mask..v3mask, vmult (nekb_set_velocity_state) and
// binvm1 (nekb_set_binv) are the COMMON state the reference reads.  i1..i3 come back dssum'ed and masked, as hmholtz leaves them.
void ophinv_(double *o1, double *o2, double *o3, double *i1, double *i2, double *i3, const double *h1, const double *h2,
             const double *tolh, const int *nmxhi)
{
    guard_fortran("ophinv", [&] {
        require_init();
        Ctx &c = ctx();
        const size_t n = (size_t)c.nelv * c.nxyz;
        NEKB_REQUIRE(c.vmask[0].n >= n && c.vmask[1].n >= n && c.vmask[2].n >= n && c.vmult.n >= n,
                     "ophinv: v1mask, v2mask, v3mask, vmult not registered (nekb_set_velocity_state)");
        NEKB_REQUIRE(c.binvm1.n >= n, "ophinv: binvm1 not registered (nekb_set_binv)");
        for (int k = 0; k < 8; k++) c.stage[k].ensure(n);
        const double *src[5] = {i1, i2, i3, h1, h2};
        for (int k = 0; k < 5; k++)
            NEKB_CUDA(cudaMemcpyAsync(c.stage[k + 3].p, src[k], n * sizeof(double), cudaMemcpyHostToDevice, c.stream));
        double *o[3] = {c.stage[0].p, c.stage[1].p, c.stage[2].p}, *r[3] = {c.stage[3].p, c.stage[4].p, c.stage[5].p};
        const double *m[3] = {c.vmask[0].p, c.vmask[1].p, c.vmask[2].p};
        ophinv_dev(o, r, c.stage[6].p, c.stage[7].p, m, c.vmult.p, c.binvm1.p, *tolh, *nmxhi, c.niter3, nullptr);
        c.niterhm = c.niter3[2];
        double *dst[3] = {o1, o2, o3}, *rdst[3] = {i1, i2, i3};
        for (int k = 0; k < 3; k++) {
            NEKB_CUDA(cudaMemcpyAsync(dst[k], o[k], n * sizeof(double), cudaMemcpyDeviceToHost, c.stream));
            NEKB_CUDA(cudaMemcpyAsync(rdst[k], r[k], n * sizeof(double), cudaMemcpyDeviceToHost, c.stream));
        }
        NEKB_CUDA(cudaStreamSynchronize(c.stream));
    });
}

// ---------------------------------------------------------------------------------------------------- gs (device API)
int nekb_gs_exchange_mode(int handle)
{
    int mode = -1;
    guard([&] {
        GsMap &h = gs_get(handle);
        mode = (ctx().nranks <= 1 || h.nshared == 0) ? 0 : (h.p2p ? 2 : 1);
    });
    return mode;
}
int nekb_gs_setup(int *handle, const int64_t *id_host, int64_t n)
{
    return guard([&] {
        require_init();
        NEKB_REQUIRE(n >= 0, "negative length");
        *handle = gs_setup_from_host_ids(id_host, n, nullptr, 0);
    });
}
int nekb_gs_setup_dev(int *handle, const int64_t *id_dev, int64_t n)
{
    return guard([&] {
        require_init();
        Ctx &c = ctx();
        if (c.nranks > 1) {  // the discovery of remote sharers runs on the host
            std::vector<int64_t> ids((size_t)n);
            NEKB_CUDA(cudaMemcpyAsync(ids.data(), id_dev, sizeof(int64_t) * n, cudaMemcpyDeviceToHost, c.stream));
            NEKB_CUDA(cudaStreamSynchronize(c.stream));
            *handle = gs_setup_from_host_ids(ids.data(), n, nullptr, 0);
            return;
        }
        const int hnd = gs_new_handle();
        gs_build_local(c.gs[hnd], id_dev, n);
        *handle = hnd;
    });
}
int nekb_gs_op_dev(int handle, double *u_dev, int op, const double *mask_dev)
{
    return guard([&] {
        require_init();
        gs_op(handle, u_dev, op, mask_dev);
    });
}
int nekb_gs_free(int handle)
{
    return guard([&] {
        gs_get(handle);
        gs_release(ctx().gs[handle]);
    });
}
int nekb_gs_info(int handle, int64_t *ngroups, int64_t *nmembers, int64_t *nshared_remote)
{
    return guard([&] {
        GsMap &h = gs_get(handle);
        if (ngroups) *ngroups = h.ngroups;
        if (nmembers) *nmembers = h.nmembers;
        if (nshared_remote) *nshared_remote = h.nshared;
    });
}
int nekb_gs_get_map(int handle, int64_t *off_host, int32_t *idx_host)
{
    return guard([&] {
        GsMap &h = gs_get(handle);
        Ctx &c = ctx();
        std::vector<int32_t> off((size_t)h.ngroups + 1, 0);
        if (h.ngroups) h.goff.download(off.data(), off.size(), c.stream);
        for (size_t i = 0; i < off.size(); i++) off_host[i] = off[i];
        if (h.nmembers) h.gidx.download(idx_host, (size_t)h.nmembers, c.stream);
    });
}
// Remote part of the map (bit-exact comparison of the exchange lists): peers[npeers], per-peer counts, and the
// concatenated ascending global ids are not kept; what is kept and returned is, per exchange item, the local
// representative index.  Sizes via nekb_gs_remote_info.
int nekb_gs_remote_info(int handle, int *npeers, int64_t *nitems)
{
    return guard([&] {
        GsMap &h = gs_get(handle);
        if (npeers) *npeers = (int)h.peers.size();
        if (nitems) *nitems = h.peer_off.empty() ? 0 : h.peer_off.back();
    });
}
int nekb_gs_get_remote(int handle, int *peers, int64_t *peer_off, int32_t *item_rep_idx)
{
    return guard([&] {
        GsMap &h = gs_get(handle);
        Ctx &c = ctx();
        for (size_t p = 0; p < h.peers.size(); p++) peers[p] = h.peers[p];
        for (size_t p = 0; p < h.peer_off.size(); p++) peer_off[p] = h.peer_off[p];
        const int64_t ni = h.peer_off.empty() ? 0 : h.peer_off.back();
        if (ni == 0) return;
        std::vector<int32_t> sid((size_t)ni), rep((size_t)h.nshared);
        h.x_item_sid.download(sid.data(), sid.size(), c.stream);
        h.x_rep.download(rep.data(), rep.size(), c.stream);
        for (int64_t q = 0; q < ni; q++) item_rep_idx[q] = rep[sid[q]];
    });
}

// ---------------------------------------------------------------------------------------------------- operators (device API)
int nekb_ax_bp5_dev(double *ap_dev, const double *p_dev, double *pap_dev)
{
    return guard([&] {
        require_init();
        launch_ax(p_dev, ap_dev, nullptr, nullptr, ctx().nelt, pap_dev);
        if (pap_dev) comm_allreduce_sum(pap_dev, 1);
    });
}
int nekb_axhelm_dev(double *au_dev, const double *u_dev, const double *h1_dev, const double *h2_dev, int imesh)
{
    return guard([&] {
        require_init();
        Ctx &c = ctx();
        launch_ax(u_dev, au_dev, h1_dev, h2_dev, imesh == 1 ? c.nelv : c.nelt, nullptr);
    });
}
int nekb_setprec_dev(double *dpc_dev, const double *h1_dev, const double *h2_dev, int imesh)
{
    return guard([&] {
        require_init();
        Ctx &c = ctx();
        setprec_run(dpc_dev, h1_dev, h2_dev, imesh == 1 ? c.nelv : c.nelt, field_handle());
    });
}
int nekb_cggos_dev(double *u_dev, const double *rhs_dev, const double *x1_dev, const double *rmult_dev, double tol,
                   int maxit, int *niter, double *hist_host)
{
    return guard([&] {
        require_init();
        Ctx &c = ctx();
        NEKB_REQUIRE(c.v1mask.n >= (size_t)field_len(), "v1mask not registered");
        CggosArgs a{u_dev, rhs_dev, x1_dev, rmult_dev, c.v1mask.p, field_handle(), c.nelt};
        const int it = cggos_run(a, tol, maxit, hist_host);
        NEKB_CUDA(cudaStreamSynchronize(c.stream));
        if (niter) *niter = it;
    });
}
int nekb_cggo_dev(double *x_dev, const double *f_dev, const double *h1_dev, const double *h2_dev, const double *mask_dev,
                  const double *mult_dev, const double *binv_dev, int imsh, double tin, int maxit, int *niter,
                  double *hist_host)
{
    return guard([&] {
        require_init();
        Ctx &c = ctx();
        const double vol = imsh == 1 ? c.volvm1 : c.voltm1;
        NEKB_REQUIRE(vol > 0.0, "volvm1/voltm1 not registered (nekb_set_step_info)");
        CggoArgs a{x_dev, f_dev, h1_dev, h2_dev, mask_dev, mult_dev, binv_dev, field_handle(), imsh == 1 ? c.nelv : c.nelt, vol, c.istep};
        c.niterhm = cggo_solve(a, tin, maxit, hist_host);
        NEKB_CUDA(cudaStreamSynchronize(c.stream));
        if (niter) *niter = c.niterhm;
    });
}

// ---------------------------------------------------------------------------------------------------- Fortran-named entry points
void fgslib_gs_setup_(int *handle, const int64_t *id, const int *n, const int *comm, const int *np)
{
    (void)comm;
    guard_fortran("fgslib_gs_setup", [&] {
        require_init();
        NEKB_REQUIRE(*np == ctx().nranks, "gs_setup: np differs from the number of ranks the library was set up for");
        *handle = gs_setup_from_host_ids(id, *n, nullptr, 0);
    });
}

static void gs_op_host(int handle, double *u, int64_t stride, int nfields, int dom, int op, int transpose)
{
    Ctx &c = ctx();
    require_init();
    NEKB_REQUIRE(dom == 1, "gs_op: only datatype 1 (double) is supported");
    NEKB_REQUIRE(transpose == 0, "gs_op: transpose must be 0");
    GsMap &h = gs_get(handle);
    DevBuf<double> &d = c.stage[0];
    d.ensure((size_t)h.n);
    for (int f = 0; f < nfields; f++) {
        double *uf = u + (int64_t)f * stride;
        NEKB_CUDA(cudaMemcpyAsync(d.p, uf, sizeof(double) * h.n, cudaMemcpyHostToDevice, c.stream));
        gs_op(handle, d.p, op, nullptr);
        NEKB_CUDA(cudaMemcpyAsync(uf, d.p, sizeof(double) * h.n, cudaMemcpyDeviceToHost, c.stream));
        NEKB_CUDA(cudaStreamSynchronize(c.stream));
    }
}
void fgslib_gs_op_(const int *handle, double *u, const int *dom, const int *op, const int *transpose)
{
    guard_fortran("fgslib_gs_op", [&] { gs_op_host(*handle, u, 0, 1, *dom, *op, *transpose); });
}
void fgslib_gs_op_many_(const int *handle, double *u1, double *u2, double *u3, double *u4, double *u5, double *u6,
                        const int *n, const int *dom, const int *op, const int *transpose)
{
    guard_fortran("fgslib_gs_op_many", [&] {
        double *us[6] = {u1, u2, u3, u4, u5, u6};
        NEKB_REQUIRE(*n >= 0 && *n <= 6, "gs_op_many: n must be 0..6");
        for (int f = 0; f < *n; f++) gs_op_host(*handle, us[f], 0, 1, *dom, *op, *transpose);
    });
}
void fgslib_gs_op_fields_(const int *handle, double *u, const int *stride, const int *n, const int *dom, const int *op,
                          const int *transpose)
{
    guard_fortran("fgslib_gs_op_fields", [&] { gs_op_host(*handle, u, *stride, *n, *dom, *op, *transpose); });
}
void fgslib_gs_free_(const int *handle)
{
    guard_fortran("fgslib_gs_free", [&] {
        gs_get(*handle);
        gs_release(ctx().gs[*handle]);
    });
}

void setupds_(int *gs_handle, const int *nx, const int *ny, const int *nz, const int *nel, const int *melg,
              const int64_t *vertex, int64_t *glo_num)
{
    (void)melg;
    guard_fortran("setupds", [&] {
        require_init();
        Ctx &c = ctx();
        NEKB_REQUIRE(*nx == *ny && *ny == *nz, "setupds: nx = ny = nz required (3-D)");
        const int64_t n = (int64_t)(*nx) * (*nx) * (*nx) * (*nel);
        setvert3d_host(glo_num, *nx, *nel, vertex, c.nranks);
        *gs_handle = gs_setup_from_host_ids(glo_num, n, nullptr, 0);
    });
}
void dssum_(double *u, const int *nx, const int *ny, const int *nz)
{
    (void)nx, (void)ny, (void)nz;
    guard_fortran("dssum", [&] { gs_op_host(field_handle(), u, 0, 1, 1, 1, 0); });
}
void dsop_(double *u, const char *op, const int *nx, const int *ny, const int *nz, size_t op_len)
{
    (void)nx, (void)ny, (void)nz;
    guard_fortran("dsop", [&] {
        char o[4] = {' ', ' ', ' ', 0};
        for (size_t i = 0; i < 3 && i < op_len; i++) o[i] = op[i];
        int code = 0;  // core/dssum.f:110-158
        if (!strcmp(o, "+  ") || !strcmp(o, "sum") || !strcmp(o, "SUM")) code = 1;
        else if (!strcmp(o, "*  ") || !strcmp(o, "mul") || !strcmp(o, "MUL")) code = 2;
        else if (!strcmp(o, "m  ") || !strcmp(o, "min") || !strcmp(o, "mna") || !strcmp(o, "MIN") || !strcmp(o, "MNA")) code = 3;
        else if (!strcmp(o, "M  ") || !strcmp(o, "max") || !strcmp(o, "mxa") || !strcmp(o, "MAX") || !strcmp(o, "MXA")) code = 4;
        NEKB_REQUIRE(code != 0, std::string("dsop: unknown operation '") + o + "'");
        gs_op_host(field_handle(), u, 0, 1, 1, code, 0);
    });
}

static int dsop_code(const char *op, size_t op_len)
{
    char o[4] = {' ', ' ', ' ', 0};
    for (size_t i = 0; i < 3 && i < op_len; i++) o[i] = op[i];
    if (!strcmp(o, "+  ") || !strcmp(o, "sum") || !strcmp(o, "SUM")) return 1;
    if (!strcmp(o, "*  ") || !strcmp(o, "mul") || !strcmp(o, "MUL")) return 2;
    if (!strcmp(o, "m  ") || !strcmp(o, "min") || !strcmp(o, "mna") || !strcmp(o, "MIN") || !strcmp(o, "MNA")) return 3;
    if (!strcmp(o, "M  ") || !strcmp(o, "max") || !strcmp(o, "mxa") || !strcmp(o, "MAX") || !strcmp(o, "MXA")) return 4;
    return 0;
}
// core/dssum.f:163-196 vec_dssum, :198-258 vec_dsop (fgslib_gs_op_many over the ldim components), :260-287 nvec_dssum
void vec_dssum_(double *u, double *v, double *w, const int *nx, const int *ny, const int *nz)
{
    (void)nx, (void)ny, (void)nz;
    guard_fortran("vec_dssum", [&] {
        double *us[3] = {u, v, w};
        for (int f = 0; f < 3; f++) gs_op_host(field_handle(), us[f], 0, 1, 1, 1, 0);
    });
}
void vec_dsop_(double *u, double *v, double *w, const int *nx, const int *ny, const int *nz, const char *op, size_t op_len)
{
    (void)nx, (void)ny, (void)nz;
    guard_fortran("vec_dsop", [&] {
        const int code = dsop_code(op, op_len);
        if (code == 0) return;  // the reference falls through silently for an unknown op (dssum.f:232-256)
        double *us[3] = {u, v, w};
        for (int f = 0; f < 3; f++) gs_op_host(field_handle(), us[f], 0, 1, 1, code, 0);
    });
}
void nvec_dssum_(double *u, const int *stride, const int *n, const int *gs_handle)
{
    guard_fortran("nvec_dssum", [&] { gs_op_host(*gs_handle, u, *stride, *n, 1, 1, 0); });
}
// core/ic.f:1871-1895 dsavg(u): u <- vmult * dssum(u) on the velocity mesh (vmult from nekb_set_velocity_state)
void dsavg_(double *u)
{
    guard_fortran("dsavg", [&] {
        require_init();
        Ctx &c = ctx();
        const size_t n = (size_t)c.nelv * c.nxyz;
        NEKB_REQUIRE(c.vmult.n >= n, "dsavg: vmult not registered (nekb_set_velocity_state)");
        c.stage[0].ensure(n);
        NEKB_CUDA(cudaMemcpyAsync(c.stage[0].p, u, n * sizeof(double), cudaMemcpyHostToDevice, c.stream));
        gs_op(field_handle(), c.stage[0].p, 1, c.vmult.p);
        NEKB_CUDA(cudaMemcpyAsync(u, c.stage[0].p, n * sizeof(double), cudaMemcpyDeviceToHost, c.stream));
        NEKB_CUDA(cudaStreamSynchronize(c.stream));
    });
}

void axhelm_(double *au, const double *u, const double *helm1, const double *helm2, const int *imesh, const int *isd)
{
    (void)isd;
    guard_fortran("axhelm", [&] {
        require_init();
        Ctx &c = ctx();
        const int nel = *imesh == 1 ? c.nelv : c.nelt;
        const size_t n = (size_t)nel * c.nxyz;
        c.stage[0].ensure(n), c.stage[1].ensure(n), c.stage[2].ensure(n), c.stage[3].ensure(n);
        NEKB_CUDA(cudaMemcpyAsync(c.stage[1].p, u, n * sizeof(double), cudaMemcpyHostToDevice, c.stream));
        NEKB_CUDA(cudaMemcpyAsync(c.stage[2].p, helm1, n * sizeof(double), cudaMemcpyHostToDevice, c.stream));
        // ifh2 as setfast decides it (core/hmholtz.f:303-305): any |h2| > 0
        bool ifh2 = false;
        for (size_t t = 0; t < n && !ifh2; t++) ifh2 = helm2[t] != 0.0;
        if (ifh2) NEKB_CUDA(cudaMemcpyAsync(c.stage[3].p, helm2, n * sizeof(double), cudaMemcpyHostToDevice, c.stream));
        launch_ax(c.stage[1].p, c.stage[0].p, c.stage[2].p, ifh2 ? c.stage[3].p : nullptr, nel, nullptr);
        NEKB_CUDA(cudaMemcpyAsync(au, c.stage[0].p, n * sizeof(double), cudaMemcpyDeviceToHost, c.stream));
        NEKB_CUDA(cudaStreamSynchronize(c.stream));
    });
}
void setprec_(double *dpcm1, const double *helm1, const double *helm2, const int *imsh, const int *isd)
{
    (void)isd;
    guard_fortran("setprec", [&] {
        require_init();
        Ctx &c = ctx();
        const int nel = *imsh == 1 ? c.nelv : c.nelt;
        const size_t n = (size_t)nel * c.nxyz;
        c.stage[0].ensure(n), c.stage[2].ensure(n), c.stage[3].ensure(n);
        NEKB_CUDA(cudaMemcpyAsync(c.stage[2].p, helm1, n * sizeof(double), cudaMemcpyHostToDevice, c.stream));
        NEKB_CUDA(cudaMemcpyAsync(c.stage[3].p, helm2, n * sizeof(double), cudaMemcpyHostToDevice, c.stream));
        setprec_run(c.stage[0].p, c.stage[2].p, c.stage[3].p, nel, field_handle());
        NEKB_CUDA(cudaMemcpyAsync(dpcm1, c.stage[0].p, n * sizeof(double), cudaMemcpyDeviceToHost, c.stream));
        NEKB_CUDA(cudaStreamSynchronize(c.stream));
    });
}
void cggo_(double *x, const double *f, const double *h1, const double *h2, const double *mask, const double *mult,
           const int *imsh, const double *tin, const int *maxit, const int *isd, const double *binv, const char *name,
           size_t name_len)
{
    (void)isd;
    guard_fortran("cggo", [&] {
        require_init();
        Ctx &c = ctx();
        if (name_len >= 4 && !strncmp(name, "PRES", 4)) {
            // hmholtz.f:641-657: ifsplit .and. name.eq.'PRES' -> x = f; hmh_gmres(x,h1,h2,mult,iter); niterhm = iter.
            // (ifsplit is implied by a completed h1mg_setup, which only the Pn-Pn pressure solver performs.)
            NEKB_REQUIRE(h1mg().ready, "cggo('PRES'): the pressure multigrid / coarse solver is not set up (nekb_h1mg_setup)");
            NEKB_REQUIRE(c.param[42] == 0.0 || c.param[42] == 1.0 || c.param[42] == 2.0, "cggo('PRES'): param(42) must be 0, 1 or 2");
            if (c.param[42] == 1.0) {   // the plain PCG below with the 'PRES' extras (coarse-grid correction, ortho)
                const size_t n = (size_t)c.nelv * c.nxyz;
                NEKB_REQUIRE(binv != nullptr || c.binvm1.n >= n, "cggo('PRES'): binvm1 not available");
                const double *src[5] = {f, h1, h2, mask, mult};
                for (int k = 0; k < 7; k++) c.stage[k].ensure(n);
                for (int k = 0; k < 5; k++)
                    NEKB_CUDA(cudaMemcpyAsync(c.stage[k + 1].p, src[k], n * sizeof(double), cudaMemcpyHostToDevice, c.stream));
                const double *d_binv = c.binvm1.p;
                if (binv != nullptr) {
                    NEKB_CUDA(cudaMemcpyAsync(c.stage[6].p, binv, n * sizeof(double), cudaMemcpyHostToDevice, c.stream));
                    d_binv = c.stage[6].p;
                }
                NEKB_REQUIRE(c.volvm1 > 0.0, "volvm1 not registered (nekb_set_step_info)");
                CggoArgs a{c.stage[0].p, c.stage[1].p, c.stage[2].p, c.stage[3].p, c.stage[4].p, c.stage[5].p, d_binv,
                           field_handle(), c.nelv, c.volvm1, c.istep};
                a.pres = true;
                c.niterhm = cggo_solve(a, *tin, *maxit, nullptr);
                NEKB_CUDA(cudaMemcpyAsync(x, c.stage[0].p, n * sizeof(double), cudaMemcpyDeviceToHost, c.stream));
                NEKB_CUDA(cudaStreamSynchronize(c.stream));
                return;
            }
            const size_t np_ = (size_t)c.nelv * c.nxyz;
            if (x != f) memcpy(x, f, np_ * sizeof(double));
            int iter = *maxit;
            if (c.param[42] == 2.0)
                hmh_flex_cg_(x, h1, h2, mult, &iter);
            else
                hmh_gmres_(x, h1, h2, mult, &iter);
            c.niterhm = iter;
            return;
        }
        const int nel = *imsh == 1 ? c.nelv : c.nelt;
        const size_t n = (size_t)nel * c.nxyz;
        const double *src[6] = {f, h1, h2, mask, mult, binv};
        for (int k = 0; k < 7; k++) c.stage[k].ensure(n);
        for (int k = 0; k < 6; k++)
            NEKB_CUDA(cudaMemcpyAsync(c.stage[k + 1].p, src[k], n * sizeof(double), cudaMemcpyHostToDevice, c.stream));
        const double vol = *imsh == 1 ? c.volvm1 : c.voltm1;
        NEKB_REQUIRE(vol > 0.0, "volvm1/voltm1 not registered (nekb_set_step_info)");
        CggoArgs a{c.stage[0].p, c.stage[1].p, c.stage[2].p, c.stage[3].p, c.stage[4].p, c.stage[5].p, c.stage[6].p,
                   field_handle(), nel, vol, c.istep};
        c.niterhm = cggo_solve(a, *tin, *maxit, nullptr);
        NEKB_CUDA(cudaMemcpyAsync(x, c.stage[0].p, n * sizeof(double), cudaMemcpyDeviceToHost, c.stream));
        NEKB_CUDA(cudaStreamSynchronize(c.stream));
    });
}
void hmholtz_(const char *name, double *u, double *rhs, const double *h1, const double *h2, const double *mask, const double *mult,
              const int *imsh, const double *tli, const int *maxit, const int *isd, size_t name_len)
{
    (void)isd;
    guard_fortran("hmholtz", [&] {
        require_init();
        Ctx &c = ctx();
        const bool pres = name_len >= 4 && !strncmp(name, "PRES", 4);
        fdm_h1_state().kfldfdm = pres ? 4 : -1;  // hmholtz.f:44-50: every call resets it (ldim+1 for 'PRES', else Jacobi)
        const int nel = *imsh == 1 ? c.nelv : c.nelt;
        const size_t n = (size_t)nel * c.nxyz;
        const DevBuf<double> &binv = *imsh == 1 ? c.binvm1 : (c.bintm1.n ? c.bintm1 : c.binvm1);
        NEKB_REQUIRE(binv.n >= n, "hmholtz: binvm1/bintm1 not registered (nekb_set_binv)");
        const double vol = *imsh == 1 ? c.volvm1 : c.voltm1;
        NEKB_REQUIRE(vol > 0.0, "volvm1/voltm1 not registered (nekb_set_step_info)");
        for (int k = 0; k < 6; k++) c.stage[k].ensure(n);
        const double *src[5] = {rhs, h1, h2, mask, mult};
        for (int k = 0; k < 5; k++)
            NEKB_CUDA(cudaMemcpyAsync(c.stage[k + 1].p, src[k], n * sizeof(double), cudaMemcpyHostToDevice, c.stream));
        double *d_rhs = c.stage[1].p;
        const double *d_h1 = c.stage[2].p, *d_h2 = c.stage[3].p, *d_mask = c.stage[4].p, *d_mult = c.stage[5].p;
        gs_op(field_handle(), d_rhs, 1, d_mask);                                   // :55-56
        double tol = fabs(*tli);
        if (c.param[22] == 0.0 || c.istep <= 10) {                                  // :59-60
            bool ifh2 = false;
            for (size_t t = 0; t < n && !ifh2; t++) ifh2 = h2[t] != 0.0;
            tol = chktcg1_dev(tol, d_rhs, d_h1, ifh2 ? d_h2 : nullptr, d_mask, d_mult, binv.p, nel, vol);
        }
        if (*tli < 0) tol = *tli;                                                  // :62
        if (pres) {  // cggo forwards 'PRES' to hmh_gmres (:641-657), which takes host arrays
            NEKB_CUDA(cudaMemcpyAsync(rhs, d_rhs, n * sizeof(double), cudaMemcpyDeviceToHost, c.stream));
            NEKB_CUDA(cudaStreamSynchronize(c.stream));
            cggo_(u, rhs, h1, h2, mask, mult, imsh, &tol, maxit, isd, nullptr, name, name_len);
            return;
        }
        CggoArgs a{c.stage[0].p, d_rhs, d_h1, d_h2, d_mask, d_mult, binv.p, field_handle(), nel, vol, c.istep};
        c.niterhm = cggo_solve(a, tol, *maxit, nullptr);
        NEKB_CUDA(cudaMemcpyAsync(u, c.stage[0].p, n * sizeof(double), cudaMemcpyDeviceToHost, c.stream));
        NEKB_CUDA(cudaMemcpyAsync(rhs, d_rhs, n * sizeof(double), cudaMemcpyDeviceToHost, c.stream));
        NEKB_CUDA(cudaStreamSynchronize(c.stream));
    });
}
void cggos_(double *u1, const double *rhs1, const double *x1, const double *rmult, const double *binv, const double *tin,
            int *maxit, const char *bpname, size_t bpname_len)
{
    (void)binv;  // bp5: dpc = 1 (setprecn, bp5.usr:300-312); binv is not read by cggos
    guard_fortran("cggos", [&] {
        require_init();
        Ctx &c = ctx();
        NEKB_REQUIRE(bpname_len >= 3 && !strncmp(bpname, "bp5", 3), "cggos: only bpname='bp5' is provided");
        NEKB_REQUIRE(c.v1mask.n >= (size_t)field_len(), "v1mask not registered");
        const size_t n = (size_t)field_len();
        for (int k = 0; k < 4; k++) c.stage[k].ensure(n);
        NEKB_CUDA(cudaMemcpyAsync(c.stage[1].p, rhs1, n * sizeof(double), cudaMemcpyHostToDevice, c.stream));
        NEKB_CUDA(cudaMemcpyAsync(c.stage[3].p, rmult, n * sizeof(double), cudaMemcpyHostToDevice, c.stream));
        if (*tin > 0.0) NEKB_CUDA(cudaMemcpyAsync(c.stage[2].p, x1, n * sizeof(double), cudaMemcpyHostToDevice, c.stream));
        CggosArgs a{c.stage[0].p, c.stage[1].p, c.stage[2].p, c.stage[3].p, c.v1mask.p, field_handle(), c.nelt};
        *maxit = cggos_run(a, *tin, *maxit, nullptr);
        NEKB_CUDA(cudaMemcpyAsync(u1, c.stage[0].p, n * sizeof(double), cudaMemcpyDeviceToHost, c.stream));
        NEKB_CUDA(cudaStreamSynchronize(c.stream));
    });
}
void axhm1_(double *pap, double *ap1, const double *p1, const double *h1, const double *h2, const char *bpname,
            size_t bpname_len)
{
    (void)h1, (void)h2;
    guard_fortran("axhm1", [&] {
        require_init();
        Ctx &c = ctx();
        NEKB_REQUIRE(bpname_len >= 3 && !strncmp(bpname, "bp5", 3), "axhm1: only bpname='bp5' is provided");
        const size_t n = (size_t)field_len();
        c.stage[0].ensure(n), c.stage[1].ensure(n);
        NEKB_CUDA(cudaMemcpyAsync(c.stage[1].p, p1, n * sizeof(double), cudaMemcpyHostToDevice, c.stream));
        double *pap_dev = &c.sc.p->work[0];
        launch_ax(c.stage[1].p, c.stage[0].p, nullptr, nullptr, c.nelt, pap_dev);
        // bp5.usr:1334-1336 accumulates the rank-local pap; the caller applies gop (bp5.usr:852)
        NEKB_CUDA(cudaMemcpyAsync(pap, pap_dev, sizeof(double), cudaMemcpyDeviceToHost, c.stream));
        NEKB_CUDA(cudaMemcpyAsync(ap1, c.stage[0].p, n * sizeof(double), cudaMemcpyDeviceToHost, c.stream));
        NEKB_CUDA(cudaStreamSynchronize(c.stream));
    });
}
double glsc3_(const double *a, const double *b, const double *mult, const int *n)
{
    double out = 0.0;
    guard_fortran("glsc3", [&] {
        require_init();
        Ctx &c = ctx();
        const size_t nn = (size_t)*n;
        c.stage[0].ensure(nn), c.stage[1].ensure(nn), c.stage[2].ensure(nn);
        NEKB_CUDA(cudaMemcpyAsync(c.stage[0].p, a, nn * sizeof(double), cudaMemcpyHostToDevice, c.stream));
        NEKB_CUDA(cudaMemcpyAsync(c.stage[1].p, b, nn * sizeof(double), cudaMemcpyHostToDevice, c.stream));
        NEKB_CUDA(cudaMemcpyAsync(c.stage[2].p, mult, nn * sizeof(double), cudaMemcpyHostToDevice, c.stream));
        double *o = &c.sc.p->work[0];
        glsc3_kernel<<<cg_grid((int64_t)nn), CG_THREADS, 0, c.stream>>>(c.stage[0].p, c.stage[1].p, c.stage[2].p, (int64_t)nn, o,
                                                                         c.partials.p, &c.sc.p->counter[0]);
        NEKB_LAUNCHED();
        comm_allreduce_sum(o, 1);
        NEKB_CUDA(cudaMemcpyAsync(&out, o, sizeof(double), cudaMemcpyDeviceToHost, c.stream));
        NEKB_CUDA(cudaStreamSynchronize(c.stream));
    });
    return out;
}

// ---------------------------------------------------------------------------------------------------- h1mg / gmres
int nekb_h1mg_setup(const int *fbc, const double *xm1, const double *ym1, const double *zm1, const int64_t *vertex, int nelv,
                    int null_space)
{
    return guard([&] {
        require_init();
        NEKB_REQUIRE(nelv >= 0 && nelv <= ctx().nelt, "h1mg_setup: nelv exceeds the registered element count");
        h1mg_setup_run(h1mg(), fbc, xm1, ym1, zm1, vertex, nelv, null_space);
    });
}
int nekb_h1mg_solve_dev(double *z_dev, double *rhs_dev)
{
    return guard([&] {
        require_init();
        h1mg_solve_dev(z_dev, rhs_dev);
    });
}
void h1mg_solve_(double *z, double *rhs, const int *if_hybrid)
{
    guard_fortran("h1mg_solve", [&] {
        require_init();
        Ctx &c = ctx();
        NEKB_REQUIRE(!(if_hybrid && *if_hybrid), "h1mg_solve: the hybrid (multiplicative) variant is not provided");
        NEKB_REQUIRE(h1mg().ready, "h1mg_solve: nekb_h1mg_setup has not been called");
        const size_t n = (size_t)h1mg().nel * c.nxyz;
        c.stage[0].ensure(n), c.stage[1].ensure(n);
        NEKB_CUDA(cudaMemcpyAsync(c.stage[1].p, rhs, n * sizeof(double), cudaMemcpyHostToDevice, c.stream));
        h1mg_solve_dev(c.stage[0].p, c.stage[1].p);
        NEKB_CUDA(cudaMemcpyAsync(z, c.stage[0].p, n * sizeof(double), cudaMemcpyDeviceToHost, c.stream));
        NEKB_CUDA(cudaMemcpyAsync(rhs, c.stage[1].p, n * sizeof(double), cudaMemcpyDeviceToHost, c.stream));  // masked in place
        NEKB_CUDA(cudaStreamSynchronize(c.stream));
    });
}
int nekb_h1mg_schwarz_dev(int level, double *e_dev, double *r_dev)
{
    return guard([&] {
        require_init();
        H1mg &M = h1mg();
        NEKB_REQUIRE(M.ready, "nekb_h1mg_setup has not been called");
        NEKB_REQUIRE(level >= 2 && level <= M.lmax, "h1mg_schwarz: level out of range (2..lmax)");
        mg_schwarz(M.lev[level - 1], r_dev, e_dev, M.nel);
    });
}
int nekb_crs_solve_dev(double *e_dev, const double *r_dev)
{
    return guard([&] {
        require_init();
        NEKB_REQUIRE(h1mg().ready, "nekb_h1mg_setup has not been called");
        crs_solve_dev(h1mg(), e_dev, r_dev);
    });
}
// ---- the reference's coarse-solver facade (core/fcrs.c:45-96), Fortran names and by-reference arguments
void crs_setup_(int *handle, const int *sid, const int *comm, const int *np, const int *n, const int64_t *id, const int *nz,
                const int *Ai, const int *Aj, const double *A, const int *null_space, const double *param,
                const char *datafname, int *ierr)
{
    (void)comm, (void)np, (void)param, (void)datafname;   // the library's own transport; XXT takes no parameters / data file
    guard_fortran("crs_setup", [&] {
        require_init();
        *handle = fcrs_setup(*sid, *n, id, *nz, Ai, Aj, A, *null_space);
        if (ierr) *ierr = 0;
    });
}
void crs_solve_(const int *handle, double *x, const double *b)
{
    guard_fortran("crs_solve", [&] { fcrs_solve_host(*handle, x, b); });
}
void crs_free_(const int *handle)
{
    guard_fortran("crs_free", [&] { fcrs_free(*handle); });
}
int nekb_fcrs_solve_dev(int handle, double *x_dev, const double *b_dev)
{
    return guard([&] { fcrs_solve_dev(handle, x_dev, b_dev); });
}
int nekb_h1mg_info(int *lmax, int *nh3, int *ntab3, int *crs_iters)
{
    return guard([&] {
        H1mg &M = h1mg();
        NEKB_REQUIRE(M.ready, "nekb_h1mg_setup has not been called");
        if (lmax) *lmax = M.lmax;
        for (int l = 0; l < M.lmax; l++) {
            if (nh3) nh3[l] = M.lev[l].nh;
            if (ntab3) ntab3[l] = M.lev[l].ntab;
        }
        if (crs_iters) {
            if (M.crs.iters_on_device && crs_scalars().p) {
                CrsScalars hs;
                NEKB_CUDA(cudaMemcpyAsync(&hs, crs_scalars().p, sizeof(CrsScalars), cudaMemcpyDeviceToHost, ctx().stream));
                NEKB_CUDA(cudaStreamSynchronize(ctx().stream));
                M.crs.last_iters = hs.it;
            }
            if (M.crs.last_iters < 0 && amg_dev().iters.p) {   // the one-launch aggregation-hierarchy CG keeps its count on the device
                NEKB_CUDA(cudaMemcpyAsync(&M.crs.last_iters, amg_dev().iters.p, sizeof(int), cudaMemcpyDeviceToHost, ctx().stream));
                NEKB_CUDA(cudaStreamSynchronize(ctx().stream));
            }
            *crs_iters = M.crs.last_iters;
        }
    });
}
int nekb_h1mg_get(const char *which, int level, double *host_out, size_t n_doubles)
{
    return guard([&] {
        H1mg &M = h1mg();
        Ctx &c = ctx();
        NEKB_REQUIRE(M.ready, "nekb_h1mg_setup has not been called");
        const std::string w(which);
        auto copy_host = [&](const std::vector<double> &v) {
            NEKB_REQUIRE(n_doubles >= v.size(), "nekb_h1mg_get: buffer too small");
            memcpy(host_out, v.data(), v.size() * sizeof(double));
        };
        if (w == "lm") return copy_host(M.lm_host);
        if (w == "ll") return copy_host(M.ll_host);
        if (w == "lr") return copy_host(M.lr_host);
        if (w == "crs_a") {
            NEKB_REQUIRE(n_doubles >= M.crs.a.n, "nekb_h1mg_get: buffer too small");
            M.crs.a.download(host_out, M.crs.a.n, c.stream);
            return;
        }
        NEKB_REQUIRE(level >= 1 && level <= M.lmax, "nekb_h1mg_get: level out of range");
        MgLevel &L = M.lev[level - 1];
        const DevBuf<double> *b = nullptr;
        if (w == "mask") b = &L.mask;
        else if (w == "rstr_wt") b = &L.rstr_wt;
        else if (w == "swt") b = &L.swt;
        else if (w == "J") b = &L.J;
        NEKB_REQUIRE(b != nullptr, std::string("nekb_h1mg_get: unknown array '") + which + "'");
        NEKB_REQUIRE(n_doubles >= b->n, "nekb_h1mg_get: buffer too small");
        b->download(host_out, b->n, c.stream);
    });
}
int nekb_crs_set_tolerance(double tol, int maxit)
{
    return guard([&] {
        NEKB_REQUIRE(tol > 0.0 && maxit > 0, "nekb_crs_set_tolerance: bad arguments");
        crs_release_graph();  // tolerance and cap are baked into the captured batch
        h1mg().crs.tol = tol;
        h1mg().crs.maxit = maxit;
    });
}
void nekb_h1mg_free(void)
{
    crs_release_graph();
    for (H1mg *M : {&h1mg(), &hsmg2()}) {
        for (MgLevel &L : M->lev) {
            if (L.gs >= 0 && L.gs < (int)ctx().gs.size()) gs_release(ctx().gs[L.gs]);
            if (L.gs_face >= 0 && L.gs_face < (int)ctx().gs.size()) gs_release(ctx().gs[L.gs_face]);
        }
        *M = H1mg();
    }
    gmres_state() = GmresState();
}

// ---------------------------------------------------------------------------------------------------- hsmg (Pn-Pn-2)
int nekb_hsmg_setup(const int *fbc, const double *xm1, const double *ym1, const double *zm1, const int64_t *vertex, int nelv,
                    int null_space, int64_t nelgv, const double *df, const double *sr, const double *ss, const double *st)
{
    return guard([&] {
        require_init();
        NEKB_REQUIRE(nelv >= 0 && nelv <= ctx().nelt, "hsmg_setup: nelv exceeds the registered element count");
        NEKB_REQUIRE((df && sr && ss && st) || (!df && !sr && !ss && !st),
                     "hsmg_setup: pass all four /fastd/ arrays df, sr, ss, st (registered) or none (computed here: gen_fast)");
        NEKB_REQUIRE(ctx().nx >= 4, "hsmg_setup: lx1 >= 4 required");
        FastdArrays f;
        f.df = df, f.sr = sr, f.ss = ss, f.st = st, f.nelgv = nelgv;
        h1mg_setup_run(hsmg2(), fbc, xm1, ym1, zm1, vertex, nelv, null_space, &f);
    });
}
// Host-only pieces of the setup, exported so that the CPU test-suite can check them against the reference without a GPU:
// the 1-D eigen-systems of gen_fast (core/fast3d.f:1351-1408) and of hsmg_setup_fast1d (core/hsmg.f:775-879).
int nekb_fast1d_sem_host(int lx1, int lbc, int rbc, double ll, double lm, double lr, double *S, double *lam)
{
    return guard([&] {
        NEKB_REQUIRE(lx1 >= 4 && lx1 <= 16, "fast1d_sem: lx1 out of range");
        std::vector<double> bh, jgl, dgl, Sv, lv;
        semhat_weighted_host(lx1 - 1, bh, jgl, dgl);
        fast1d_sem_host(lbc, rbc, ll, lm, lr, bh, jgl, dgl, Sv, lv);
        memcpy(S, Sv.data(), sizeof(double) * (size_t)lx1 * lx1);
        memcpy(lam, lv.data(), sizeof(double) * (size_t)lx1);
    });
}
int nekb_fast1d_host(int n, int lbc, int rbc, double ll, double lm, double lr, double *S, double *lam)
{
    return guard([&] {
        NEKB_REQUIRE(n >= 1 && n <= 16, "fast1d: polynomial order out of range");
        std::vector<double> ah, bh, zh, Sv, lv;
        semhat_host(n, ah, bh, zh);
        fast1d_host(lbc, rbc, ll, lm, lr, ah, bh, n, Sv, lv);
        const int nl = n + 3;
        memcpy(S, Sv.data(), sizeof(double) * (size_t)nl * nl);
        memcpy(lam, lv.data(), sizeof(double) * (size_t)nl);
    });
}
int nekb_hsmg_solve_dev(double *e_dev, const double *r_dev)
{
    return guard([&] {
        require_init();
        hsmg_solve_dev(e_dev, r_dev);
    });
}
int nekb_local_solves_fdm_dev(double *u_dev, const double *v_dev)
{
    return guard([&] {
        require_init();
        local_solves_fdm_dev(u_dev, v_dev);
    });
}
static void pnpn2_host(const char *who, double *out, const double *in, bool whole)
{
    guard_fortran(who, [&] {
        require_init();
        Ctx &c = ctx();
        H1mg &M = hsmg2();
        NEKB_REQUIRE(M.ready, "nekb_hsmg_setup has not been called");
        const size_t n = (size_t)M.lev[M.lmax - 1].n;
        c.stage[0].ensure(n), c.stage[1].ensure(n);
        NEKB_CUDA(cudaMemcpyAsync(c.stage[1].p, in, n * sizeof(double), cudaMemcpyHostToDevice, c.stream));
        if (whole)
            hsmg_solve_dev(c.stage[0].p, c.stage[1].p);
        else
            local_solves_fdm_dev(c.stage[0].p, c.stage[1].p);
        NEKB_CUDA(cudaMemcpyAsync(out, c.stage[0].p, n * sizeof(double), cudaMemcpyDeviceToHost, c.stream));
        NEKB_CUDA(cudaStreamSynchronize(c.stream));
    });
}
void hsmg_solve_(double *e, const double *r) { pnpn2_host("hsmg_solve", e, r, true); }
void local_solves_fdm_(double *u, const double *v) { pnpn2_host("local_solves_fdm", u, v, false); }
int nekb_hsmg_get(const char *which, int level, double *host_out, size_t n_doubles)
{
    return guard([&] {
        H1mg &M = hsmg2();
        Ctx &c = ctx();
        NEKB_REQUIRE(M.ready, "nekb_hsmg_setup has not been called");
        NEKB_REQUIRE(level >= 1 && level <= M.lmax, "nekb_hsmg_get: level out of range");
        MgLevel &L = M.lev[level - 1];
        const std::string w(which);
        const DevBuf<double> *b = nullptr;
        if (w == "J") b = &L.J;
        else if (w == "owt") b = &L.owt;
        else if (w == "swt") b = &L.swt;
        else if (w == "mask") b = &L.mask;
        NEKB_REQUIRE(b != nullptr && b->n > 0, std::string("nekb_hsmg_get: unknown or empty array '") + which + "'");
        NEKB_REQUIRE(n_doubles >= b->n, "nekb_hsmg_get: buffer too small");
        b->download(host_out, b->n, c.stream);
    });
}

int nekb_fdm_h1_setup(const int *face_internal, const double *mask, const double *xm1, const double *ym1, const double *zm1, int nel)
{
    return guard([&] {
        require_init();
        fdm_h1_setup(face_internal, mask, xm1, ym1, zm1, nel);
    });
}
int nekb_set_kfldfdm(int kfldfdm)
{
    return guard([&] { fdm_h1_state().kfldfdm = kfldfdm; });
}
int nekb_set_fdm_prec_h1b_dev(double *d_dev, const double *h1_dev, const double *h2_dev)
{
    return guard([&] {
        require_init();
        set_fdm_prec_h1b_dev(d_dev, h1_dev, h2_dev, fdm_h1_state().nel);
    });
}
int nekb_fdm_h1_dev(double *z_dev, const double *r_dev, const double *d_dev, const double *mask_dev)
{
    return guard([&] {
        require_init();
        fdm_h1_apply(z_dev, r_dev, d_dev, mask_dev, fdm_h1_state().nel, field_handle());
    });
}
void set_fdm_prec_h1b_(double *d, const double *h1, const double *h2, const int *nel)
{
    guard_fortran("set_fdm_prec_h1b", [&] {
        require_init();
        Ctx &c = ctx();
        const size_t n = (size_t)(*nel) * c.nxyz;
        for (int k = 0; k < 3; k++) c.stage[k].ensure(n);
        NEKB_CUDA(cudaMemcpyAsync(c.stage[1].p, h1, n * sizeof(double), cudaMemcpyHostToDevice, c.stream));
        NEKB_CUDA(cudaMemcpyAsync(c.stage[2].p, h2, n * sizeof(double), cudaMemcpyHostToDevice, c.stream));
        set_fdm_prec_h1b_dev(c.stage[0].p, c.stage[1].p, c.stage[2].p, *nel);
        NEKB_CUDA(cudaMemcpyAsync(d, c.stage[0].p, n * sizeof(double), cudaMemcpyDeviceToHost, c.stream));
        NEKB_CUDA(cudaStreamSynchronize(c.stream));
    });
}
void fdm_h1_(double *z, const double *r, const double *d, const double *mask, const double *mult, const int *nel, const int *kt,
             double *rr)
{
    (void)mult, (void)kt;
    guard_fortran("fdm_h1", [&] {
        require_init();
        Ctx &c = ctx();
        const size_t n = (size_t)(*nel) * c.nxyz;
        for (int k = 0; k < 4; k++) c.stage[k].ensure(n);
        NEKB_CUDA(cudaMemcpyAsync(c.stage[1].p, r, n * sizeof(double), cudaMemcpyHostToDevice, c.stream));
        NEKB_CUDA(cudaMemcpyAsync(c.stage[2].p, d, n * sizeof(double), cudaMemcpyHostToDevice, c.stream));
        NEKB_CUDA(cudaMemcpyAsync(c.stage[3].p, mask, n * sizeof(double), cudaMemcpyHostToDevice, c.stream));
        fdm_h1_apply(c.stage[0].p, c.stage[1].p, c.stage[2].p, c.stage[3].p, *nel, field_handle());
        NEKB_CUDA(cudaMemcpyAsync(z, c.stage[0].p, n * sizeof(double), cudaMemcpyDeviceToHost, c.stream));
        NEKB_CUDA(cudaStreamSynchronize(c.stream));
        if (rr) memcpy(rr, r, n * sizeof(double));  // hmholtz.f:965 rr = r (ifbhalf = .false.)
    });
}
int nekb_fdm_h1_get(const char *which, void *host_out, size_t n_bytes)
{
    return guard([&] {
        FdmH1State &F = fdm_h1_state();
        NEKB_REQUIRE(F.ready, "nekb_fdm_h1_setup has not been called");
        const std::string w(which);
        if (w == "ktype") {
            NEKB_REQUIRE(n_bytes >= F.ktype_host.size() * sizeof(int32_t), "nekb_fdm_h1_get: buffer too small");
            int32_t *o = (int32_t *)host_out;
            for (size_t i = 0; i < F.ktype_host.size(); i++) o[i] = F.ktype_host[i] + 1;
        } else if (w == "elsize") {
            NEKB_REQUIRE(n_bytes >= F.elsize_host.size() * sizeof(double), "nekb_fdm_h1_get: buffer too small");
            memcpy(host_out, F.elsize_host.data(), F.elsize_host.size() * sizeof(double));
        } else if (w == "dd") {
            NEKB_REQUIRE(n_bytes >= F.dd_host.size() * sizeof(double), "nekb_fdm_h1_get: buffer too small");
            memcpy(host_out, F.dd_host.data(), F.dd_host.size() * sizeof(double));
        } else
            NEKB_REQUIRE(false, std::string("nekb_fdm_h1_get: unknown array '") + which + "'");
    });
}

int nekb_set_pressure_state(const double *pmask, const double *binvm1, double tolps, double param21, int ifvcor, int64_t nelgv)
{
    return guard([&] {
        require_init();
        Ctx &c = ctx();
        GmresState &G = gmres_state();
        const size_t n = (size_t)c.nelv * c.nxyz;
        if (pmask) G.pmask.upload(pmask, n, c.stream);
        if (binvm1) G.binvm1.upload(binvm1, n, c.stream);
        G.tolps = tolps, G.param21 = param21, G.ifvcor = ifvcor;
        G.ntotg = (double)nelgv * (double)c.nxyz;
        NEKB_CUDA(cudaStreamSynchronize(c.stream));
    });
}
int nekb_hmh_gmres_dev(double *res_dev, const double *h1_dev, const double *h2_dev, const double *wt_dev, const double *pmask_dev,
                       double tol, int maxit, int *iter, double *hist_host, double *div0)
{
    return guard([&] {
        require_init();
        Ctx &c = ctx();
        NEKB_REQUIRE(h1mg().ready, "hmh_gmres: nekb_h1mg_setup has not been called");
        NEKB_REQUIRE(c.volvm1 > 0.0, "volvm1 not registered (nekb_set_step_info)");
        const int it = hmh_gmres_run(res_dev, h1_dev, h2_dev, wt_dev, pmask_dev, h1mg().nel, h1mg().lev[h1mg().lmax - 1].gs,
                                     c.volvm1, tol, maxit, hist_host, div0);
        if (iter) *iter = it;
    });
}
// gmres.f:338-342 (tolerance guard, param(21) / istep overrides) + the solve, on device pointers; h2 == nullptr: h2 = 0
static int hmh_gmres_body(double *res, const double *h1, const double *h2, const double *wt, int maxit)
{
    Ctx &c = ctx();
    GmresState &G = gmres_state();
    H1mg &M = h1mg();
    NEKB_REQUIRE(M.ready, "hmh_gmres: nekb_h1mg_setup has not been called");
    NEKB_REQUIRE(c.volvm1 > 0.0, "volvm1 not registered (nekb_set_step_info)");
    const size_t n = (size_t)M.nel * c.nxyz;
    NEKB_REQUIRE(G.pmask.n >= n && G.binvm1.n >= n, "hmh_gmres: pmask/binvm1 not registered (nekb_set_pressure_state)");
    double tolps = chktcg1_dev(G.tolps, res, h1, h2, G.pmask.p, wt, G.binvm1.p, M.nel, c.volvm1);
    if (G.param21 > 0 && tolps > fabs(G.param21)) tolps = fabs(G.param21);
    if (c.istep == 0) tolps = 1.e-4;
    const double tol = G.param21 < 0 ? -fabs(G.param21) : tolps;
    return hmh_gmres_run(res, h1, h2, wt, G.pmask.p, M.nel, M.lev[M.lmax - 1].gs, c.volvm1, tol, maxit, nullptr, nullptr);
}
// core/hmholtz.f:2164-2290 hmh_flex_cg(res,h1,h2,wt,iter): flexible PCG with h1mg_solve (param(42) = 2) on device pointers
static int hmh_flex_cg_body(double *res, const double *h1, const double *h2, const double *wt, int maxit)
{
    Ctx &c = ctx();
    GmresState &G = gmres_state();
    H1mg &M = h1mg();
    NEKB_REQUIRE(M.ready, "hmh_flex_cg: nekb_h1mg_setup has not been called");
    NEKB_REQUIRE(c.volvm1 > 0.0, "volvm1 not registered (nekb_set_step_info)");
    const int64_t n = (int64_t)M.nel * c.nxyz;
    NEKB_REQUIRE(G.pmask.n >= (size_t)n && G.binvm1.n >= (size_t)n, "hmh_flex_cg: pmask/binvm1 not registered (nekb_set_pressure_state)");
    cudaStream_t s = c.stream;
    const int grid = cg_grid(n), gsh = M.lev[M.lmax - 1].gs;
    double tolps = chktcg1_dev(G.tolps, res, h1, h2, G.pmask.p, wt, G.binvm1.p, M.nel, c.volvm1);   // :2204-2208
    if (G.param21 > 0 && tolps > fabs(G.param21)) tolps = fabs(G.param21);
    if (c.istep == 0) tolps = 1.e-4;
    double tolpss = tolps;
    static DevBuf<double> r, r1, p, z, w;
    r.ensure((size_t)n), r1.ensure((size_t)n), p.ensure((size_t)n), z.ensure((size_t)n), w.ensure((size_t)n);
    NEKB_CUDA(cudaMemcpyAsync(r.p, res, sizeof(double) * (size_t)n, cudaMemcpyDeviceToDevice, s));
    NEKB_CUDA(cudaMemsetAsync(r1.p, 0, sizeof(double) * (size_t)n, s));
    NEKB_CUDA(cudaMemsetAsync(p.p, 0, sizeof(double) * (size_t)n, s));
    NEKB_CUDA(cudaMemsetAsync(res, 0, sizeof(double) * (size_t)n, s));
    double rho1 = 1.0;
    const double div0 = sqrt(gm_glsc3(r.p, wt, r.p, n) / c.volvm1);
    if (G.param21 < 0) tolpss = fabs(G.param21) * div0;
    int iter = 0;
    for (int k = 1; k <= maxit; k++) {
        h1mg_solve_dev(z.p, r.p);                                               // z = M^-1 r (masks r in place, as the reference)
        gm_sub3_kernel<<<grid, 256, 0, s>>>(r1.p, r1.p, r.p, n);                // sub2(r1,r)
        NEKB_LAUNCHED();
        const double rho0 = rho1;
        rho1 = gm_glsc3(z.p, wt, r.p, n);
        const double rho2 = -gm_glsc3(z.p, wt, r1.p, n);
        const double beta = rho2 / rho0;
        NEKB_CUDA(cudaMemcpyAsync(r1.p, r.p, sizeof(double) * (size_t)n, cudaMemcpyDeviceToDevice, s));
        gm_cmult2_kernel<<<grid, 256, 0, s>>>(p.p, p.p, beta, n);               // add2s1(p,z,beta): p = beta p + z
        NEKB_LAUNCHED();
        proj_axpy(p.p, z.p, 1.0, n);
        gm_ax(w.p, p.p, h1, h2, G.pmask.p, M.nel, gsh);                          // w = A p
        const double den = gm_glsc3(w.p, wt, p.p, n);
        const double alpha = rho1 / den;
        proj_axpy(res, p.p, alpha, n);
        proj_axpy(r.p, w.p, -alpha, n);
        const double rnorm = sqrt(gm_glsc3(r.p, r.p, wt, n) / c.volvm1);
        iter++;
        if (rnorm < tolpss) break;
    }
    gm_ortho(res, n);
    NEKB_CUDA(cudaStreamSynchronize(s));
    return iter;
}
void hmh_flex_cg_(double *res, const double *h1, const double *h2, const double *wt, int *iter)
{
    guard_fortran("hmh_flex_cg", [&] {
        require_init();
        Ctx &c = ctx();
        H1mg &M = h1mg();
        NEKB_REQUIRE(M.ready, "hmh_flex_cg: nekb_h1mg_setup has not been called");
        const size_t n = (size_t)M.nel * c.nxyz;
        for (int k = 0; k < 4; k++) c.stage[k].ensure(n);
        const double *src[4] = {res, h1, h2, wt};
        for (int k = 0; k < 4; k++)
            NEKB_CUDA(cudaMemcpyAsync(c.stage[k].p, src[k], n * sizeof(double), cudaMemcpyHostToDevice, c.stream));
        bool ifh2 = false;
        for (size_t t = 0; t < n && !ifh2; t++) ifh2 = h2[t] != 0.0;
        *iter = hmh_flex_cg_body(c.stage[0].p, c.stage[1].p, ifh2 ? c.stage[2].p : nullptr, c.stage[3].p, *iter);
        NEKB_CUDA(cudaMemcpyAsync(res, c.stage[0].p, n * sizeof(double), cudaMemcpyDeviceToHost, c.stream));
        NEKB_CUDA(cudaStreamSynchronize(c.stream));
    });
}
void hmh_gmres_(double *res, const double *h1, const double *h2, const double *wt, int *iter)
{
    guard_fortran("hmh_gmres", [&] {
        require_init();
        Ctx &c = ctx();
        H1mg &M = h1mg();
        NEKB_REQUIRE(M.ready, "hmh_gmres: nekb_h1mg_setup has not been called");
        const size_t n = (size_t)M.nel * c.nxyz;
        for (int k = 0; k < 4; k++) c.stage[k].ensure(n);
        const double *src[4] = {res, h1, h2, wt};
        for (int k = 0; k < 4; k++)
            NEKB_CUDA(cudaMemcpyAsync(c.stage[k].p, src[k], n * sizeof(double), cudaMemcpyHostToDevice, c.stream));
        bool ifh2 = false;
        for (size_t t = 0; t < n && !ifh2; t++) ifh2 = h2[t] != 0.0;
        *iter = hmh_gmres_body(c.stage[0].p, c.stage[1].p, ifh2 ? c.stage[2].p : nullptr, c.stage[3].p, *iter);
        NEKB_CUDA(cudaMemcpyAsync(res, c.stage[0].p, n * sizeof(double), cudaMemcpyDeviceToHost, c.stream));
        NEKB_CUDA(cudaStreamSynchronize(c.stream));
    });
}

// ---------------------------------------------------------------------------------------------------- Pn-Pn-2 E operator
// core/navier1.f:258-293 cdabdtp(ap,wp,h1,h2,h2inv,intype) on device pointers
static void cdabdtp_dev(double *ap, const double *wp, const double *h1, const double *h2, const double *h2inv, int intype)
{
    Ctx &c = ctx();
    const size_t n = (size_t)c.nelv * c.nxyz;
    Pnpn2Work &W = pnpn2_work();
    for (int k = 0; k < 3; k++) W.ta[k].ensure(n), W.tb[k].ensure(n);
    opgradt_dev(W.ta[0].p, W.ta[1].p, W.ta[2].p, wp);
    if (intype == 0 || intype == -1) {   // D (h1 A + h2 B)^-1 D^T: the three velocity solves (tolhs, nmxv)
        NEKB_REQUIRE(c.vmask[0].n >= n && c.vmask[1].n >= n && c.vmask[2].n >= n && c.vmult.n >= n && c.binvm1.n >= n,
                     "cdabdtp(intype 0/-1): v1mask..v3mask, vmult, binvm1 not registered");
        double *o[3] = {W.tb[0].p, W.tb[1].p, W.tb[2].p}, *r[3] = {W.ta[0].p, W.ta[1].p, W.ta[2].p};
        const double *m[3] = {c.vmask[0].p, c.vmask[1].p, c.vmask[2].p};
        ophinv_dev(o, r, h1, h2, m, c.vmult.p, c.binvm1.p, mesh2().tolhs, mesh2().nmxv, c.niter3, nullptr);
    } else
        opbinv_dev(W.tb[0].p, W.tb[1].p, W.tb[2].p, W.ta[0].p, W.ta[1].p, W.ta[2].p, h2inv, field_handle());
    opdiv_dev(ap, W.tb[0].p, W.tb[1].p, W.tb[2].p);
}
// core/navier1.f:1089-1154 chktcg2
static double chktcg2_dev(double tol, const double *res, int64_t n2)
{
    Ctx &c = ctx();
    Mesh2 &M = mesh2();
    const double eps = M.prelax != 0.0 ? M.prelax : 1.e-10;
    M.t1.ensure((size_t)n2);
    uz_resid_kernel<<<cg_grid(n2), 256, 0, c.stream>>>(M.t1.p, M.bm2inv.p, res, nullptr, n2);   // ta = res*bm2inv
    NEKB_LAUNCHED();
    const double rinit = sqrt(gm_glsc3(M.t1.p, M.t1.p, M.bm2.p, n2) / M.volvm2);
    if (rinit < tol) return tol;
    const double rmin = M.tolpdf > 0.0 ? M.tolpdf : eps * rinit;
    if (tol < rmin) tol = rmin;
    if (M.ifvcor) {
        const double otr = gm_glsc3(res, M.ones.p, M.ones.p, n2);
        const double tolmin = fabs(otr) * 100.0;
        if (tol < tolmin) tol = tolmin;
    }
    return tol;
}
// core/gmres.f:2-237 uzawa_gmres(res,h1,h2,h2inv,intype,iter) on device pointers: right-preconditioned GMRES(lgmres) on
// E = cdabdtp with hsmg_solve as preconditioner, split weights ml = sqrt(bm2inv), mu = sqrt(bm2).  Returns iter.
static int uzawa_gmres_dev(double *res, const double *h1, const double *h2, const double *h2inv, int intype)
{
    Ctx &c = ctx();
    Mesh2 &M = mesh2();
    require_mesh2();
    cudaStream_t s = c.stream;
    const int64_t n2 = (int64_t)c.nelv * 216;
    NEKB_REQUIRE(M.bm2.n >= (size_t)n2 && M.bm2inv.n >= (size_t)n2 && M.volvm2 > 0.0, "uzawa_gmres: bm2, bm2inv, volvm2 not registered");
    const int m = gmres_state().m, grid = cg_grid(n2);
    GmresState &G = gmres_state();
    G.scal.ensure(64);
    c.partials.ensure((size_t)(GM_NV > 4 ? GM_NV : 4) * CG_PART_STRIDE);
    if ((int)M.V.size() != m + 1) M.V.resize(m + 1), M.Z.resize(m);
    M.r.ensure((size_t)n2), M.w.ensure((size_t)n2), M.x.ensure((size_t)n2);
    if (M.ones.n < (size_t)n2) {
        M.ones.alloc((size_t)n2);
        fill_kernel<<<grid, 256, 0, s>>>(M.ones.p, 1.0, n2);
        NEKB_LAUNCHED();
    }
    if (M.ml.n < (size_t)n2) {
        M.ml.alloc((size_t)n2), M.mu.alloc((size_t)n2);
        uz_split_kernel<<<grid, 256, 0, s>>>(M.ml.p, M.mu.p, M.bm2.p, M.bm2inv.p, n2);
        NEKB_LAUNCHED();
    }
    auto Vj = [&](int j) -> double * {
        M.V[j].ensure((size_t)n2);
        return M.V[j].p;
    };
    auto Zj = [&](int j) -> double * {
        M.Z[j].ensure((size_t)n2);
        return M.Z[j].p;
    };
    const double norm_fac = 1.0 / sqrt(M.volvm2);
    double tolps = chktcg2_dev(M.tolps, res, n2);
    if (M.param21 > 0 && tolps > fabs(M.param21)) tolps = fabs(M.param21);
    if (c.istep == 0) tolps = 1.e-4;
    double tolpss = tolps;
    std::vector<double> H((size_t)(m + 1) * m, 0.0), cg(m, 0.0), sg(m, 0.0), gam(m + 1, 0.0), cvec(m, 0.0), hcol(m + 1);
    auto Hm = [&](int i, int j) -> double & { return H[(size_t)i * m + j]; };
    NEKB_CUDA(cudaMemsetAsync(M.x.p, 0, sizeof(double) * (size_t)n2, s));
    int iter = 0, j = 0;
    bool conv = false;
    double div0 = 0.0, rnorm = 0.0;
    unsigned *counter = &c.sc.p->counter[0];
    const double *wt = M.ones.p;
    while (!conv && iter < 100) {
        if (iter == 0)
            uz_resid_kernel<<<grid, 256, 0, s>>>(M.r.p, M.ml.p, res, nullptr, n2);
        else {
            cdabdtp_dev(M.w.p, M.x.p, h1, h2, h2inv, intype);
            uz_resid_kernel<<<grid, 256, 0, s>>>(M.r.p, M.ml.p, res, M.w.p, n2);
        }
        NEKB_LAUNCHED();
        gam[0] = sqrt(gm_glsc3(M.r.p, M.r.p, wt, n2));
        if (iter == 0) {
            div0 = gam[0] * norm_fac;
            if (M.param21 < 0) tolpss = fabs(M.param21) * div0;
        }
        rnorm = 0.0;
        if (gam[0] == 0.0) break;
        gm_cmult2_kernel<<<grid, 256, 0, s>>>(Vj(0), M.r.p, 1.0 / gam[0], n2);
        NEKB_LAUNCHED();
        for (j = 0; j < m; j++) {
            iter++;
            uz_resid_kernel<<<grid, 256, 0, s>>>(M.w.p, M.mu.p, Vj(j), nullptr, n2);     // w = U^-1 v_j
            NEKB_LAUNCHED();
            hsmg_solve_dev(Zj(j), M.w.p);                                                 // z_j = M^-1 w
            cdabdtp_dev(M.w.p, Zj(j), h1, h2, h2inv, intype);                             // w = E z_j
            col2_kernel<<<grid, 256, 0, s>>>(M.w.p, M.ml.p, n2);                          // w = L^-1 w
            NEKB_LAUNCHED();
            for (int i0 = 0; i0 <= j; i0 += GM_NV) {
                const int nv = (j + 1 - i0) < GM_NV ? (j + 1 - i0) : GM_NV;
                GmPtrs P;
                for (int q = 0; q < GM_NV; q++) P.v[q] = Vj(i0 + (q < nv ? q : 0)), P.h[q] = 0.0;
                gm_dots_kernel<<<grid, 256, 0, s>>>(M.w.p, wt, P, nv, n2, G.scal.p + i0, c.partials.p, counter);
                NEKB_LAUNCHED();
            }
            gm_reduce_to_host(G.scal.p, j + 1, hcol.data());
            for (int i = 0; i <= j; i++) Hm(i, j) = hcol[i];
            double alpha2 = 0.0;
            for (int i0 = 0; i0 <= j; i0 += GM_NV) {
                const int nv = (j + 1 - i0) < GM_NV ? (j + 1 - i0) : GM_NV;
                const bool last = i0 + GM_NV > j;
                GmPtrs P;
                for (int q = 0; q < GM_NV; q++) P.v[q] = Vj(i0 + (q < nv ? q : 0)), P.h[q] = q < nv ? hcol[i0 + q] : 0.0;
                gm_project_kernel<<<grid, 256, 0, s>>>(M.w.p, wt, P, nv, n2, last ? 1 : 0, G.scal.p + 40, c.partials.p, counter);
                NEKB_LAUNCHED();
            }
            gm_reduce_to_host(G.scal.p + 40, 1, &alpha2);
            for (int i = 0; i < j; i++) {
                const double t = Hm(i, j);
                Hm(i, j) = cg[i] * t + sg[i] * Hm(i + 1, j);
                Hm(i + 1, j) = -sg[i] * t + cg[i] * Hm(i + 1, j);
            }
            const double alpha = sqrt(alpha2);
            rnorm = 0.0;
            if (alpha == 0.0) {
                conv = true;
                break;
            }
            const double l = sqrt(Hm(j, j) * Hm(j, j) + alpha * alpha), t = 1.0 / l;
            cg[j] = Hm(j, j) * t;
            sg[j] = alpha * t;
            Hm(j, j) = l;
            gam[j + 1] = -sg[j] * gam[j];
            gam[j] = cg[j] * gam[j];
            rnorm = fabs(gam[j + 1]) * norm_fac;
            if (rnorm < tolpss) {
                conv = true;
                break;
            }
            if (j == m - 1) break;  // restart
            gm_cmult2_kernel<<<grid, 256, 0, s>>>(Vj(j + 1), M.w.p, 1.0 / alpha, n2);
            NEKB_LAUNCHED();
        }
        const int kk = j + 1 > m ? m : j + 1;
        for (int k = kk - 1; k >= 0; k--) {
            double t = gam[k];
            for (int i = kk - 1; i > k; i--) t = t - Hm(k, i) * cvec[i];
            cvec[k] = t / Hm(k, k);
        }
        for (int i0 = 0; i0 < kk; i0 += GM_NV) {
            const int nv = (kk - i0) < GM_NV ? (kk - i0) : GM_NV;
            GmPtrs P;
            for (int q = 0; q < GM_NV; q++) P.v[q] = Zj(i0 + (q < nv ? q : 0)), P.h[q] = q < nv ? cvec[i0 + q] : 0.0;
            gm_combine_kernel<<<grid, 256, 0, s>>>(M.x.p, P, nv, n2);
            NEKB_LAUNCHED();
        }
    }
    M.divex = rnorm, M.div0 = div0;
    NEKB_CUDA(cudaMemcpyAsync(res, M.x.p, sizeof(double) * (size_t)n2, cudaMemcpyDeviceToDevice, s));
    if (M.ifvcor) {  // ortho (core/navier1.f:223-257)
        const double sum = gm_glsc3(res, M.ones.p, M.ones.p, n2);
        gm_cadd_kernel<<<grid, 256, 0, s>>>(res, -sum / ((double)M.nelgv * 216.0), n2);
        NEKB_LAUNCHED();
    }
    NEKB_CUDA(cudaStreamSynchronize(s));
    return iter;
}
int nekb_set_uzawa_state(double tolps, double param21, double prelax, double tolpdf)
{
    return guard([&] {
        Mesh2 &M = mesh2();
        M.tolps = tolps, M.param21 = param21, M.prelax = prelax, M.tolpdf = tolpdf;
    });
}
int nekb_uzawa_gmres_dev(double *res, const double *h1, const double *h2, const double *h2inv, int intype, int *iter, double *div0,
                         double *divex)
{
    return guard([&] {
        require_init();
        const int it = uzawa_gmres_dev(res, h1, h2, h2inv, intype);
        if (iter) *iter = it;
        if (div0) *div0 = mesh2().div0;
        if (divex) *divex = mesh2().divex;
    });
}
void uzawa_gmres_(double *res, const double *h1, const double *h2, const double *h2inv, const int *intype, int *iter)
{
    guard_fortran("uzawa_gmres", [&] {
        require_init();
        Ctx &c = ctx();
        const size_t n = (size_t)c.nelv * c.nxyz, n2 = (size_t)c.nelv * 216;
        for (int k = 0; k < 4; k++) c.stage[k].ensure(n);
        NEKB_CUDA(cudaMemcpyAsync(c.stage[0].p, res, n2 * sizeof(double), cudaMemcpyHostToDevice, c.stream));
        const double *src[3] = {h1, h2, h2inv};
        for (int k = 0; k < 3; k++) NEKB_CUDA(cudaMemcpyAsync(c.stage[k + 1].p, src[k], n * sizeof(double), cudaMemcpyHostToDevice, c.stream));
        *iter = uzawa_gmres_dev(c.stage[0].p, c.stage[1].p, c.stage[2].p, c.stage[3].p, *intype);
        NEKB_CUDA(cudaMemcpyAsync(res, c.stage[0].p, n2 * sizeof(double), cudaMemcpyDeviceToHost, c.stream));
        NEKB_CUDA(cudaStreamSynchronize(c.stream));
    });
}
int nekb_set_mesh2(int lx2, const double *ixm12, const double *dxm12, const double *w3m2, const double *const *metrics9, const double *bm2,
                   const double *bm2inv, double volvm2, double tolhs, int nmxv, int64_t nelgv, int ifvcor)
{
    return guard([&] {
        require_init();
        Ctx &c = ctx();
        Mesh2 &M = mesh2();
        NEKB_REQUIRE(lx2 == c.nx - 2, "set_mesh2: lx2 must be lx1 - 2");
        const int lx1 = c.nx;
        // ixm12(lx2,lx1), dxm12(lx2,lx1) arrive column-major (Fortran): element (a,i) at a + lx2*i
        std::vector<double> I((size_t)lx2 * lx1), D((size_t)lx2 * lx1);
        for (int a = 0; a < lx2; a++)
            for (int i = 0; i < lx1; i++) I[(size_t)a * lx1 + i] = ixm12[a + lx2 * i], D[(size_t)a * lx1 + i] = dxm12[a + lx2 * i];
        M.i12.upload(I.data(), I.size(), c.stream), M.d12.upload(D.data(), D.size(), c.stream);
        NEKB_REQUIRE(lx2 == 6 && lx1 == 8, "set_mesh2: the Pn-Pn-2 kernels are built for lx1 = 8, lx2 = 6");
        NEKB_CUDA(cudaMemcpyToSymbolAsync(c_I12, I.data(), 48 * sizeof(double), 0, cudaMemcpyHostToDevice, c.stream));
        NEKB_CUDA(cudaMemcpyToSymbolAsync(c_D12, D.data(), 48 * sizeof(double), 0, cudaMemcpyHostToDevice, c.stream));
        const size_t p2 = (size_t)lx2 * lx2 * lx2, n2 = p2 * c.nelv;
        M.w3.upload(w3m2, p2, c.stream);
        for (int k = 0; k < 9; k++) M.met[k].upload(metrics9[k], n2, c.stream);
        if (bm2) M.bm2.upload(bm2, n2, c.stream);
        if (bm2inv) M.bm2inv.upload(bm2inv, n2, c.stream);
        M.lx2 = lx2, M.volvm2 = volvm2, M.tolhs = tolhs, M.nmxv = nmxv, M.nelgv = nelgv, M.ifvcor = ifvcor != 0;
        M.ml.release(), M.mu.release();
        NEKB_CUDA(cudaStreamSynchronize(c.stream));
        M.ready = true;
    });
}
int nekb_opgradt_dev(double *ox, double *oy, double *oz, const double *p)
{
    return guard([&] {
        require_init();
        opgradt_dev(ox, oy, oz, p);
    });
}
int nekb_opdiv_dev(double *out, const double *ux, const double *uy, const double *uz)
{
    return guard([&] {
        require_init();
        opdiv_dev(out, ux, uy, uz);
    });
}
int nekb_cdabdtp_dev(double *ap, const double *wp, const double *h1, const double *h2, const double *h2inv, int intype)
{
    return guard([&] {
        require_init();
        cdabdtp_dev(ap, wp, h1, h2, h2inv, intype);
    });
}
void opgradt_(double *outx, double *outy, double *outz, const double *inpfld)
{
    guard_fortran("opgradt", [&] {
        require_init();
        Ctx &c = ctx();
        require_mesh2();
        const size_t n = (size_t)c.nelv * c.nxyz, n2 = (size_t)c.nelv * 216;
        for (int k = 0; k < 4; k++) c.stage[k].ensure(n);
        NEKB_CUDA(cudaMemcpyAsync(c.stage[3].p, inpfld, n2 * sizeof(double), cudaMemcpyHostToDevice, c.stream));
        opgradt_dev(c.stage[0].p, c.stage[1].p, c.stage[2].p, c.stage[3].p);
        double *dst[3] = {outx, outy, outz};
        for (int k = 0; k < 3; k++) NEKB_CUDA(cudaMemcpyAsync(dst[k], c.stage[k].p, n * sizeof(double), cudaMemcpyDeviceToHost, c.stream));
        NEKB_CUDA(cudaStreamSynchronize(c.stream));
    });
}
void opdiv_(double *outfld, const double *inpx, const double *inpy, const double *inpz)
{
    guard_fortran("opdiv", [&] {
        require_init();
        Ctx &c = ctx();
        require_mesh2();
        const size_t n = (size_t)c.nelv * c.nxyz, n2 = (size_t)c.nelv * 216;
        for (int k = 0; k < 4; k++) c.stage[k].ensure(n);
        const double *src[3] = {inpx, inpy, inpz};
        for (int k = 0; k < 3; k++) NEKB_CUDA(cudaMemcpyAsync(c.stage[k].p, src[k], n * sizeof(double), cudaMemcpyHostToDevice, c.stream));
        opdiv_dev(c.stage[3].p, c.stage[0].p, c.stage[1].p, c.stage[2].p);
        NEKB_CUDA(cudaMemcpyAsync(outfld, c.stage[3].p, n2 * sizeof(double), cudaMemcpyDeviceToHost, c.stream));
        NEKB_CUDA(cudaStreamSynchronize(c.stream));
    });
}
void opbinv_(double *out1, double *out2, double *out3, double *inp1, double *inp2, double *inp3, const double *h2inv)
{
    guard_fortran("opbinv", [&] {
        require_init();
        Ctx &c = ctx();
        const size_t n = (size_t)c.nelv * c.nxyz;
        for (int k = 0; k < 7; k++) c.stage[k].ensure(n);
        const double *src[4] = {inp1, inp2, inp3, h2inv};
        for (int k = 0; k < 4; k++) NEKB_CUDA(cudaMemcpyAsync(c.stage[k + 3].p, src[k], n * sizeof(double), cudaMemcpyHostToDevice, c.stream));
        opbinv_dev(c.stage[0].p, c.stage[1].p, c.stage[2].p, c.stage[3].p, c.stage[4].p, c.stage[5].p, c.stage[6].p, field_handle());
        double *dst[6] = {out1, out2, out3, inp1, inp2, inp3};
        for (int k = 0; k < 6; k++) NEKB_CUDA(cudaMemcpyAsync(dst[k], c.stage[k].p, n * sizeof(double), cudaMemcpyDeviceToHost, c.stream));
        NEKB_CUDA(cudaStreamSynchronize(c.stream));
    });
}
void cdabdtp_(double *ap, const double *wp, const double *h1, const double *h2, const double *h2inv, const int *intype)
{
    guard_fortran("cdabdtp", [&] {
        require_init();
        Ctx &c = ctx();
        require_mesh2();
        const size_t n = (size_t)c.nelv * c.nxyz, n2 = (size_t)c.nelv * 216;
        for (int k = 0; k < 5; k++) c.stage[k].ensure(n);
        NEKB_CUDA(cudaMemcpyAsync(c.stage[1].p, wp, n2 * sizeof(double), cudaMemcpyHostToDevice, c.stream));
        const double *src[3] = {h1, h2, h2inv};
        for (int k = 0; k < 3; k++) NEKB_CUDA(cudaMemcpyAsync(c.stage[k + 2].p, src[k], n * sizeof(double), cudaMemcpyHostToDevice, c.stream));
        cdabdtp_dev(c.stage[0].p, c.stage[1].p, c.stage[2].p, c.stage[3].p, c.stage[4].p, *intype);
        NEKB_CUDA(cudaMemcpyAsync(ap, c.stage[0].p, n2 * sizeof(double), cudaMemcpyDeviceToHost, c.stream));
        NEKB_CUDA(cudaStreamSynchronize(c.stream));
    });
}

// ---------------------------------------------------------------------------------------------------- hsolve + projection
int nekb_set_projection(int ifield, int ifprojfld, int ldimt_proj)
{
    return guard([&] {
        require_init();
        NEKB_REQUIRE(ifield >= 0 && ifield < 16, "set_projection: ifield out of range");
        ctx().ifprojfld[ifield] = ifprojfld != 0;
        if (ldimt_proj >= 0) ctx().ldimt_proj = ldimt_proj;
    });
}
int nekb_projection_reset(void)
{
    return guard([&] { proj_states().clear(); });
}
// navier4.f:562-634 on device pointers.  Returns niterhm.  napprox (host, may be NULL): ivar(1) = mmx, ivar(2) = m.
static int hsolve_dev(const char *name, size_t name_len, double *u, double *r, const double *h1, const double *h2, const double *vmk,
                      const double *vml, int imsh, double tol, int maxit, const double *bi, int *napprox)
{
    Ctx &c = ctx();
    char cname[5] = "    ";
    for (size_t k = 0; k < 4 && k < name_len; k++) cname[k] = (char)toupper((unsigned char)name[k]);
    const bool pres = !strncmp(cname, "PRES", 4);
    const int nel = imsh == 1 ? c.nelv : c.nelt;
    const int64_t n = (int64_t)nel * c.nxyz;
    const double vol = imsh == 1 ? c.volvm1 : c.voltm1;
    NEKB_REQUIRE(vol > 0.0, "volvm1/voltm1 not registered (nekb_set_step_info)");
    const DevBuf<double> &binvc = imsh == 1 ? c.binvm1 : (c.bintm1.n ? c.bintm1 : c.binvm1);
    const double *binv_chk = binvc.n >= (size_t)n ? binvc.p : bi;   // hmholtz / hmhzpf read binvm1 from COMMON
    // ifh2
    absmax_kernel<<<cg_grid(n), CG_THREADS, 0, c.stream>>>(h2, n, &c.sc.p->work[3], c.partials.p, &c.sc.p->counter[0]);
    NEKB_LAUNCHED();
    comm_allreduce_max(&c.sc.p->work[3], 1);
    double h2max = 0.0;
    NEKB_CUDA(cudaMemcpyAsync(&h2max, &c.sc.p->work[3], sizeof(double), cudaMemcpyDeviceToHost, c.stream));
    NEKB_CUDA(cudaStreamSynchronize(c.stream));
    const bool ifh2 = h2max > 0.0;

    bool ifstdh = true;                                                        // :588-601
    if (c.ifprojfld[c.ifield < 16 ? c.ifield : 0]) ifstdh = false;
    double p945 = c.param[94];
    if (pres) ifstdh = false, p945 = c.param[95];
    if (c.ifield > c.ldimt_proj + 1) ifstdh = true;
    if (c.param[93] == 0.0) ifstdh = true;
    if (p945 == 0.0) ifstdh = true;
    if ((double)c.istep < p945) ifstdh = true;

    auto solve = [&](double tolin, bool through_hmholtz) -> int {
        double t = tolin;
        if (through_hmholtz) {                                                 // hmholtz.f:55-62
            gs_op(field_handle(), r, 1, vmk);
            t = fabs(tolin);
            if (c.param[22] == 0.0 || c.istep <= 10) t = chktcg1_dev(t, r, h1, ifh2 ? h2 : nullptr, vmk, vml, binv_chk, nel, vol);
            if (tolin < 0) t = tolin;
        } else {                                                               // hmhzpf, navier4.f:536-538
            if (c.param[22] != 0.0) t = fabs(c.param[22]);
            t = chktcg1_dev(t, r, h1, ifh2 ? h2 : nullptr, vmk, vml, binv_chk, nel, vol);
        }
        if (pres && c.param[42] == 1.0) {                                      // cggo :660-846 with the 'PRES' extras
            NEKB_REQUIRE(h1mg().ready, "hsolve('PRES'): needs nekb_h1mg_setup (the coarse solver of crs_solve_h1)");
            CggoArgs a{u, r, h1, h2, vmk, vml, bi, field_handle(), nel, vol, c.istep};
            a.pres = true;
            // hmholtz.f:50 (hsolve reaches cggo through hmholtz / hmhzpf): kfldfdm = ldim + 1 for 'PRES'; the next hmholtz call
            // of the reference resets it, so the value registered before is put back for whatever solve comes next
            const int kf_prev = fdm_h1_state().kfldfdm;
            fdm_h1_state().kfldfdm = 4;
            const int it_pcg = cggo_solve(a, t, maxit, nullptr);
            fdm_h1_state().kfldfdm = kf_prev;
            return it_pcg;
        }
        if (pres) {                                                            // cggo :641-657
            NEKB_REQUIRE(h1mg().ready && (c.param[42] == 0.0 || c.param[42] == 2.0),
                         "hsolve('PRES'): needs nekb_h1mg_setup and param(42) = 0 (GMRES), 1 (PCG) or 2 (flexible CG)");
            NEKB_CUDA(cudaMemcpyAsync(u, r, sizeof(double) * (size_t)n, cudaMemcpyDeviceToDevice, c.stream));
            if (c.param[42] == 2.0) return hmh_flex_cg_body(u, h1, ifh2 ? h2 : nullptr, vml, maxit);
            return hmh_gmres_body(u, h1, ifh2 ? h2 : nullptr, vml, maxit);
        }
        CggoArgs a{u, r, h1, h2, vmk, vml, bi, field_handle(), nel, vol, c.istep};
        return cggo_solve(a, t, maxit, nullptr);
    };
    if (ifstdh) return solve(tol, true);

    col2_kernel<<<cg_grid(n), 256, 0, c.stream>>>(r, vmk, n);                  // :607-608
    NEKB_LAUNCHED();
    gs_op(field_handle(), r, 1, nullptr);
    ProjState &S = proj_get(std::string(cname, 4), n, 20);
    if (napprox && napprox[1] >= 0 && napprox[1] < S.m) S.m = napprox[1];      // the caller restarted the space
    project1_dev(S, r, h1, h2, vmk, vml, nel, field_handle(), ifh2);
    const int it = solve(tol, false);
    project2_dev(S, u, h1, h2, vmk, vml, nel, field_handle(), ifh2);
    if (napprox) napprox[0] = S.mmx, napprox[1] = S.m;
    return it;
}
int nekb_hsolve_dev(const char *name4, double *u, double *r, const double *h1, const double *h2, const double *vmk, const double *vml,
                    int imsh, double tol, int maxit, const double *bi, int *napprox, int *niter)
{
    return guard([&] {
        require_init();
        const int it = hsolve_dev(name4, strlen(name4), u, r, h1, h2, vmk, vml, imsh, tol, maxit, bi, napprox);
        ctx().niterhm = it;
        if (niter) *niter = it;
    });
}
void hsolve_(const char *name, double *u, double *r, const double *h1, const double *h2, const double *vmk, const double *vml,
             const int *imsh, const double *tol, const int *maxit, const int *isd, double *approx, int *napprox, const double *bi,
             size_t name_len)
{
    (void)isd, (void)approx;  // the approximation space lives on the device (proj.cuh); ivar(1:2) are mirrored into napprox
    guard_fortran("hsolve", [&] {
        require_init();
        Ctx &c = ctx();
        const int nel = *imsh == 1 ? c.nelv : c.nelt;
        const size_t n = (size_t)nel * c.nxyz;
        for (int k = 0; k < 7; k++) c.stage[k].ensure(n);
        const double *src[6] = {r, h1, h2, vmk, vml, bi};
        for (int k = 0; k < 6; k++)
            NEKB_CUDA(cudaMemcpyAsync(c.stage[k + 1].p, src[k], n * sizeof(double), cudaMemcpyHostToDevice, c.stream));
        c.niterhm = hsolve_dev(name, name_len, c.stage[0].p, c.stage[1].p, c.stage[2].p, c.stage[3].p, c.stage[4].p, c.stage[5].p, *imsh,
                               *tol, *maxit, c.stage[6].p, napprox);
        NEKB_CUDA(cudaMemcpyAsync(u, c.stage[0].p, n * sizeof(double), cudaMemcpyDeviceToHost, c.stream));
        NEKB_CUDA(cudaMemcpyAsync(r, c.stage[1].p, n * sizeof(double), cudaMemcpyDeviceToHost, c.stream));
        NEKB_CUDA(cudaStreamSynchronize(c.stream));
    });
}

// ---------------------------------------------------------------------------------------------------- host-side setup
// ---------------------------------------------------------------------------------------------------- mesh files
int nekb_re2_info(const char *path, int64_t *nelgt, int *ldim, int64_t *nelgv, int *wdsize, int64_t *ncurve, int *nsections,
                  int64_t *nbc, int nbc_cap)
{
    return guard([&] {
        FileHandle fh(path);
        NEKB_REQUIRE(fh.f != nullptr, std::string("Cannot open .re2 file! ") + path);
        const Re2Header h = re2_header(fh);
        const Re2Sections s = re2_sections(fh, h);
        if (nelgt) *nelgt = h.nelgt;
        if (ldim) *ldim = h.ldim;
        if (nelgv) *nelgv = h.nelgv;
        if (wdsize) *wdsize = h.wdsize;
        if (ncurve) *ncurve = s.ncurve;
        if (nsections) *nsections = (int)s.nbc.size();
        for (int k = 0; nbc && k < nbc_cap && k < (int)s.nbc.size(); k++) nbc[k] = s.nbc[k];
    });
}
int nekb_re2_read_mesh(const char *path, int64_t e0, int64_t nel, double *xc, double *yc, double *zc, int *igroup)
{
    return guard([&] {
        FileHandle fh(path);
        NEKB_REQUIRE(fh.f != nullptr, std::string("Cannot open .re2 file! ") + path);
        const Re2Header h = re2_header(fh);
        NEKB_REQUIRE(xc && yc && (zc || h.ldim == 2), "re2_read_mesh: output arrays required");
        re2_read_mesh(fh, h, e0, nel, xc, yc, zc, igroup);
    });
}
int nekb_re2_read_bc(const char *path, int section, char *cbc, double *bc)
{
    return guard([&] {
        FileHandle fh(path);
        NEKB_REQUIRE(fh.f != nullptr, std::string("Cannot open .re2 file! ") + path);
        const Re2Header h = re2_header(fh);
        const Re2Sections s = re2_sections(fh, h);
        re2_read_bc(fh, h, s, section, cbc, bc);
    });
}
int nekb_re2_read_curves(const char *path, char *ccurve, double *curve)
{
    return guard([&] {
        FileHandle fh(path);
        NEKB_REQUIRE(fh.f != nullptr, std::string("Cannot open .re2 file! ") + path);
        const Re2Header h = re2_header(fh);
        const Re2Sections s = re2_sections(fh, h);
        re2_read_curves(fh, h, s, ccurve, curve);
    });
}
int nekb_ma2_info(const char *path, int64_t *nel, int64_t *hdr7)
{
    return guard([&] {
        FileHandle fh(path);
        NEKB_REQUIRE(fh.f != nullptr, std::string("Cannot find map file! ") + path);
        const Ma2Header h = ma2_header(fh);
        if (nel) *nel = h.nel;
        if (hdr7) {
            const int64_t v[7] = {h.nel, h.nactive, h.depth, h.d2, h.npts, h.nrank, h.noutflow};
            for (int k = 0; k < 7; k++) hdr7[k] = v[k];
        }
    });
}
int nekb_ma2_read(const char *path, int nlv, int64_t e0, int64_t nel, int32_t *leaf, int64_t *vertex)
{
    return guard([&] {
        FileHandle fh(path);
        NEKB_REQUIRE(fh.f != nullptr, std::string("Cannot find map file! ") + path);
        NEKB_REQUIRE(nlv == 4 || nlv == 8, "ma2_read: nlv must be 4 or 8");
        const Ma2Header h = ma2_header(fh);
        ma2_read(fh, h, nlv, e0, nel, leaf, vertex);
    });
}
int nekb_co2_info(const char *path, int64_t *nelgt, int64_t *nelgv, int *nv)
{
    return guard([&] {
        FileHandle fh(path);
        NEKB_REQUIRE(fh.f != nullptr, std::string("Cannot find con file! ") + path);
        const Co2Header h = co2_header(fh);
        if (nelgt) *nelgt = h.nelgt;
        if (nelgv) *nelgv = h.nelgv;
        if (nv) *nv = h.nv;
    });
}
int nekb_co2_read(const char *path, int nlv, int64_t e0, int64_t nel, int64_t *eid, int64_t *vertex)
{
    return guard([&] {
        FileHandle fh(path);
        NEKB_REQUIRE(fh.f != nullptr, std::string("Cannot find con file! ") + path);
        const Co2Header h = co2_header(fh);
        NEKB_REQUIRE(nlv == h.nv, "Number of vertices do not match!");          // map2.f:443-444
        co2_read(fh, h, e0, nel, eid, vertex);
    });
}
// ---- host set-up of the aggregation hierarchy for large coarse problems (crs_amg.cuh; no device work)
int nekb_crs_amg_build_host(int64_t n, int64_t nz, const int64_t *I, const int64_t *J, const double *V, int64_t nmax, double theta,
                            double omega_p, int *nlevels)
{
    return guard([&] {
        NEKB_REQUIRE(n >= 1 && nz >= 1 && I && J && V, "crs_amg_build_host: bad arguments");
        std::vector<int64_t> vi(I, I + nz), vj(J, J + nz);
        std::vector<double> vv(V, V + nz);
        amg_host_hierarchy() = amg_build(csr_from_triplets(n, vi, vj, vv), nmax, theta, omega_p);
        if (nlevels) *nlevels = (int)amg_host_hierarchy().A.size();
    });
}
int nekb_crs_amg_level_info(int level, int64_t *n, int64_t *nnz)
{
    return guard([&] {
        AmgHierarchy &H = amg_host_hierarchy();
        NEKB_REQUIRE(level >= 0 && level < (int)H.A.size(), "crs_amg_level_info: no such level");
        if (n) *n = H.A[level].n;
        if (nnz) *nnz = H.A[level].nnz();
    });
}
int nekb_crs_amg_level_get(int level, int64_t *rowptr, int32_t *col, double *val, int32_t *agg)
{
    return guard([&] {
        AmgHierarchy &H = amg_host_hierarchy();
        NEKB_REQUIRE(level >= 0 && level < (int)H.A.size(), "crs_amg_level_get: no such level");
        const CsrHost &A = H.A[level];
        if (rowptr) std::copy(A.rowptr.begin(), A.rowptr.end(), rowptr);
        if (col) std::copy(A.col.begin(), A.col.end(), col);
        if (val) std::copy(A.val.begin(), A.val.end(), val);
        if (agg) {
            NEKB_REQUIRE(level < (int)H.agg.size(), "crs_amg_level_get: the coarsest level has no aggregates");
            std::copy(H.agg[level].begin(), H.agg[level].end(), agg);
        }
    });
}
int nekb_crs_amg_level_p(int level, int64_t *nnz, int64_t *rowptr, int32_t *col, double *val)
{
    return guard([&] {
        AmgHierarchy &H = amg_host_hierarchy();
        NEKB_REQUIRE(level >= 0 && level < (int)H.P.size(), "crs_amg_level_p: no prolongation at this level");
        const CsrHost &P = H.P[level];
        if (nnz) *nnz = P.nnz();
        if (rowptr) std::copy(P.rowptr.begin(), P.rowptr.end(), rowptr);
        if (col) std::copy(P.col.begin(), P.col.end(), col);
        if (val) std::copy(P.val.begin(), P.val.end(), val);
    });
}
int nekb_crs_amg_upload(double omega)
{
    return guard([&] {
        require_init();
        NEKB_REQUIRE(omega > 0.0 && omega <= 1.0, "crs_amg_upload: damping outside (0, 1]");
        amg_upload(omega);
    });
}
int nekb_crs_amg_solve_dev(double *x_dev, const double *b_dev, double tol, int maxit, int *iters)
{
    return guard([&] {
        require_init();
        const int it = (amg_coop_enabled() && !amg_dev().L.empty()) ? amg_pcg_solve_coop(x_dev, b_dev, tol, maxit, true)
                                                                      : amg_pcg_solve(x_dev, b_dev, tol, maxit);
        if (iters) *iters = it;
    });
}
int nekb_assign_gllnid(int *gllnid, int64_t nelgt, int64_t nelgv, int np)
{
    return guard([&] {
        NEKB_REQUIRE(gllnid && nelgt >= 0, "assign_gllnid: bad arguments");
        std::vector<int> g(gllnid, gllnid + nelgt);
        assign_gllnid(g, nelgt, nelgv, np);
        memcpy(gllnid, g.data(), (size_t)nelgt * sizeof(int));
    });
}
int nekb_setvert3d(int64_t *glo_num, int64_t *ngv, int nx, int64_t nel, const int64_t *vertex, int np)
{
    return guard([&] {
        NEKB_REQUIRE(nx >= 2 && nel >= 0, "setvert3d: bad sizes");
        const int64_t v = setvert3d_host(glo_num, nx, nel, vertex, np);
        if (ngv) *ngv = v;
    });
}
// Host-only discovery of the ids shared with other ranks (uses the registered transport).  ids: this rank's
// distinct non-zero ids, ascending.  Two-call protocol: with out arrays NULL only the sizes are returned.
int nekb_gs_discover(const int64_t *uniq_ids, int64_t n, int *npeers, int64_t *nitems, int *peers, int64_t *peer_off,
                     int64_t *item_ids)
{
    return guard([&] {
        std::vector<int64_t> u(uniq_ids, uniq_ids + n);
        static SharedIds cache;
        if (peers == nullptr) cache = discover_shared_ids(u);
        int64_t tot = 0;
        for (auto &v : cache.ids) tot += (int64_t)v.size();
        if (npeers) *npeers = (int)cache.peers.size();
        if (nitems) *nitems = tot;
        if (peers == nullptr) return;
        int64_t o = 0;
        for (size_t p = 0; p < cache.peers.size(); p++) {
            peers[p] = cache.peers[p];
            peer_off[p] = o;
            for (int64_t id : cache.ids[p]) item_ids[o++] = id;
        }
        peer_off[cache.peers.size()] = o;
    });
}

int nekb_bp5_setup(int nelx, int nely, int nelz, int px, int py, int pz, double deform)
{
    return guard([&] {
        require_init();
        bp5_setup(nelx, nely, nelz, px, py, pz, deform);
    });
}
int nekb_bp5_solve(double tol, int maxit, int *niter, double *seconds, double *hist_host)
{
    return guard([&] {
        require_init();
        Ctx &c = ctx();
        Bp5Case &b = bp5case();
        NEKB_REQUIRE(b.built, "nekb_bp5_setup has not been called");
        CggosArgs a{b.u1.p, b.r1.p, b.e1.p, b.mult.p, b.mask.p, b.gs_handle, (int)b.nel};
        NEKB_CUDA(cudaEventRecord(c.ev0, c.stream));
        const int it = cggos_run(a, tol, maxit, hist_host);
        NEKB_CUDA(cudaEventRecord(c.ev1, c.stream));
        NEKB_CUDA(cudaEventSynchronize(c.ev1));
        float ms = 0.f;
        NEKB_CUDA(cudaEventElapsedTime(&ms, c.ev0, c.ev1));
        if (niter) *niter = it;
        if (seconds) *seconds = 1e-3 * ms;
    });
}
int nekb_bp5_relerr(double *relerr)
{
    return guard([&] {
        require_init();
        Ctx &c = ctx();
        Bp5Case &b = bp5case();
        NEKB_REQUIRE(b.built, "nekb_bp5_setup has not been called");
        double *o = &c.sc.p->work[0];
        glrdif_kernel<<<cg_grid(b.n), CG_THREADS, 0, c.stream>>>(b.u1.p, b.e1.p, b.n, o, c.partials.p, &c.sc.p->counter[0]);
        NEKB_LAUNCHED();
        comm_allreduce_max(o, 3);
        double v[3];
        NEKB_CUDA(cudaMemcpyAsync(v, o, sizeof v, cudaMemcpyDeviceToHost, c.stream));
        NEKB_CUDA(cudaStreamSynchronize(c.stream));
        const double xm = v[1] > v[2] ? v[1] : v[2];  // bp5.usr:422-447 glrdif
        *relerr = xm > 0 ? v[0] / xm : -v[0];
    });
}
static void *bp5_ptr(const char *which, size_t *bytes)
{
    Ctx &c = ctx();
    Bp5Case &b = bp5case();
    NEKB_REQUIRE(b.built, "nekb_bp5_setup has not been called");
    const size_t nb = (size_t)b.n * sizeof(double);
    const std::string w(which);
    *bytes = nb;
    if (w == "u1") return b.u1.p;
    if (w == "e1") return b.e1.p;
    if (w == "r1") return b.r1.p;
    if (w == "mask") return b.mask.p;
    if (w == "mult") return b.mult.p;
    if (w == "xm1") return b.xm1.p;
    if (w == "ym1") return b.ym1.p;
    if (w == "zm1") return b.zm1.p;
    if (w == "bm1") return c.bm1.p;
    if (w == "glo_num") {
        *bytes = (size_t)b.n * sizeof(int64_t);
        return b.glo_num.p;
    }
    if (w == "g") {  // device layout [e][6][nxyz]
        *bytes = 6 * nb;
        return c.g.p;
    }
    NEKB_REQUIRE(false, "unknown array name '" + w + "'");
    return nullptr;
}
static void *bp5_ptr_checked(const char *which, size_t *bytes)
{
    void *p = bp5_ptr(which, bytes);
    NEKB_REQUIRE(p != nullptr, std::string("array '") + which + "' was released (lean BP5 set-up above 10^6 elements, NEKB_BP5_LEAN)");
    return p;
}
int nekb_bp5_get(const char *which, void *host_out, size_t n_bytes)
{
    return guard([&] {
        require_init();
        Ctx &c = ctx();
        if (!strcmp(which, "gf")) {  // reference layout gf(6,nxyz,nel)
            const int64_t total = (int64_t)6 * bp5case().n;
            NEKB_REQUIRE(n_bytes >= (size_t)total * sizeof(double), "output buffer too small");
            DevBuf<double> tmp;
            tmp.alloc((size_t)total);
            gf_interleave_kernel<<<blocks_for(total), 256, 0, c.stream>>>(tmp.p, c.g.p, c.nxyz, total);
            NEKB_LAUNCHED();
            tmp.download((double *)host_out, (size_t)total, c.stream);
            return;
        }
        size_t bytes = 0;
        void *p = bp5_ptr_checked(which, &bytes);
        NEKB_REQUIRE(n_bytes >= bytes, "output buffer too small");
        NEKB_CUDA(cudaMemcpyAsync(host_out, p, bytes, cudaMemcpyDeviceToHost, c.stream));
        NEKB_CUDA(cudaStreamSynchronize(c.stream));
    });
}
int64_t nekb_bp5_nel_local(void) { return bp5case().nel; }
int nekb_bp5_gs_handle(void) { return bp5case().gs_handle; }
void *nekb_bp5_devptr(const char *which)
{
    void *p = nullptr;
    guard([&] {
        size_t bytes;
        p = bp5_ptr_checked(which, &bytes);
    });
    return p;
}

// Plain device memory helpers for hosts without a CUDA runtime binding of their own (ctypes tests).
void *nekb_dev_alloc(size_t bytes)
{
    void *p = nullptr;
    if (guard([&] { NEKB_CUDA(cudaMalloc(&p, bytes ? bytes : 1)); })) return nullptr;
    return p;
}
void nekb_dev_free(void *p) { cudaFree(p); }
int nekb_h2d(void *dev, const void *host, size_t bytes)
{
    return guard([&] {
        NEKB_CUDA(cudaMemcpyAsync(dev, host, bytes, cudaMemcpyHostToDevice, ctx().stream));
        NEKB_CUDA(cudaStreamSynchronize(ctx().stream));
    });
}
int nekb_d2h(void *host, const void *dev, size_t bytes)
{
    return guard([&] {
        NEKB_CUDA(cudaMemcpyAsync(host, dev, bytes, cudaMemcpyDeviceToHost, ctx().stream));
        NEKB_CUDA(cudaStreamSynchronize(ctx().stream));
    });
}
int nekb_d2d(void *dst_dev, const void *src_dev, size_t bytes)
{
    return guard([&] {
        NEKB_CUDA(cudaMemcpyAsync(dst_dev, src_dev, bytes, cudaMemcpyDeviceToDevice, ctx().stream));
        NEKB_CUDA(cudaStreamSynchronize(ctx().stream));
    });
}
int nekb_sync(void)
{
    return guard([&] { NEKB_CUDA(cudaStreamSynchronize(ctx().stream)); });
}

}  // extern "C"
