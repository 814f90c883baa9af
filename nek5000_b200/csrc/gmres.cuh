// gmres.cuh -- right-preconditioned GMRES(m) for the pressure Helmholtz problem.
//
// Replaces core/gmres.f:304-545 hmh_gmres (with `ax`, :284-302, and the tolerance guard chktcg1,
// core/hmholtz.f:527-609).  ml = mu = 1 (uzawa_gmres_split, gmres.f:252-271); the preconditioner is h1mg_solve
// (ifmgrid, param(40) in 0..2); one-pass classical Gram-Schmidt with a single reduction per column (:417-426).
//
// Device design: the Krylov bases V (m+1 vectors) and Z (m vectors) stay in HBM (2*30+1 fields, as the reference's
// core/GMRES:10-23).  The j+1 inner products of a column are computed by multi-vector kernels that read w and the
// weight once per 8 basis vectors, the projection w -= sum h_i v_i is fused the same way and its last pass also
// produces (w,w).  The small Hessenberg / Givens recurrences run on the host exactly as written in the reference
// (one device->host copy of j+1 numbers per column, where the reference has an MPI_Allreduce).
#pragma once
#include "fdm_h1.cuh"

namespace nekb {

constexpr int GM_NV = 8;
struct GmPtrs {
    const double *v[GM_NV];
    double h[GM_NV];
};

// out[i] = sum w * v_i * wt, i < nv   (vlsc3, core/navier4.f:323, for nv basis vectors at once)
__global__ void __launch_bounds__(256)
    gm_dots_kernel(const double *__restrict__ w, const double *__restrict__ wt, GmPtrs P, int nv, int64_t n, double *out,
                   double *partials, unsigned *counter)
{
    __shared__ double red[33];
    __shared__ int s_last;
    double s[GM_NV];
#pragma unroll
    for (int i = 0; i < GM_NV; i++) s[i] = 0.0;
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < n; t += (int64_t)gridDim.x * blockDim.x) {
        const double ww = w[t] * wt[t];
#pragma unroll
        for (int i = 0; i < GM_NV; i++)
            if (i < nv) s[i] = fma(ww, P.v[i][t], s[i]);
    }
#pragma unroll
    for (int i = 0; i < GM_NV; i++) {
        const double b = block_reduce(s[i], red);
        if (threadIdx.x == 0) partials[(size_t)i * gridDim.x + blockIdx.x] = b;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        const unsigned t = atomicInc(counter, gridDim.x - 1);
        s_last = (t == gridDim.x - 1);
    }
    __syncthreads();
    if (s_last) {
        __threadfence();
        for (int i = 0; i < nv; i++) {
            double a = 0.0;
            for (unsigned q = threadIdx.x; q < gridDim.x; q += blockDim.x) a += __ldcg(partials + (size_t)i * gridDim.x + q);
            const double tot = block_reduce(a, red);
            if (threadIdx.x == 0) out[i] = tot;
        }
    }
}

// w -= sum_i h_i v_i ; optionally out = sum w*w*wt   (add2s2 loop gmres.f:424-426 + glsc3 :448)
__global__ void __launch_bounds__(256)
    gm_project_kernel(double *__restrict__ w, const double *__restrict__ wt, GmPtrs P, int nv, int64_t n, int want_norm,
                      double *out, double *partials, unsigned *counter)
{
    __shared__ double red[33];
    double s = 0.0;
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < n; t += (int64_t)gridDim.x * blockDim.x) {
        double v = w[t];
#pragma unroll
        for (int i = 0; i < GM_NV; i++)
            if (i < nv) v = fma(-P.h[i], P.v[i][t], v);
        w[t] = v;
        if (want_norm) s = fma(v * v, wt[t], s);
    }
    if (want_norm) {
        const double b = block_reduce(s, red);
        grid_reduce(b, partials, counter, red, [=](double tot) { *out = tot; });
    }
}

// x += sum_i c_i z_i   (gmres.f:518-520)
__global__ void __launch_bounds__(256) gm_combine_kernel(double *__restrict__ x, GmPtrs P, int nv, int64_t n)
{
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < n; t += (int64_t)gridDim.x * blockDim.x) {
        double v = x[t];
#pragma unroll
        for (int i = 0; i < GM_NV; i++)
            if (i < nv) v = fma(P.h[i], P.v[i][t], v);
        x[t] = v;
    }
}
// a = s * b
__global__ void __launch_bounds__(256) gm_cmult2_kernel(double *__restrict__ a, const double *__restrict__ b, double s, int64_t n)
{
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < n; t += (int64_t)gridDim.x * blockDim.x) a[t] = b[t] * s;
}
// a = b - c
__global__ void __launch_bounds__(256)
    gm_sub3_kernel(double *__restrict__ a, const double *__restrict__ b, const double *__restrict__ c, int64_t n)
{
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < n; t += (int64_t)gridDim.x * blockDim.x) a[t] = b[t] - c[t];
}
// out = sum a [* b [* c]]
__global__ void __launch_bounds__(256)
    gm_sum3_kernel(const double *__restrict__ a, const double *__restrict__ b, const double *__restrict__ c, int64_t n, double *out,
                   double *partials, unsigned *counter)
{
    __shared__ double red[33];
    double s = 0.0;
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < n; t += (int64_t)gridDim.x * blockDim.x) {
        double v = a[t];
        if (b) v *= b[t];
        if (c) v *= c[t];
        s += v;
    }
    const double bs = block_reduce(s, red);
    grid_reduce(bs, partials, counter, red, [=](double tot) { *out = tot; });
}
__global__ void __launch_bounds__(256) gm_cadd_kernel(double *__restrict__ a, double s, int64_t n)
{
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < n; t += (int64_t)gridDim.x * blockDim.x) a[t] += s;
}

struct GmresState {
    int m = 30;  // lgmres, core/SIZE.template:31
    std::vector<DevBuf<double>> V, Z;
    DevBuf<double> w, r, x, scal;
    // registered pressure-solve state (nekb_set_pressure_state)
    DevBuf<double> pmask, binvm1;
    double tolps = 1e-8, param21 = 0.0;
    int ifvcor = 0;
    double ntotg = 0.0;  // lx1^3 * nelgv
};
inline GmresState &gmres_state()
{
    static GmresState g;
    return g;
}

inline double gm_reduce_to_host(double *dev, int count, double *host)
{
    Ctx &c = ctx();
    comm_allreduce_sum(dev, count);
    NEKB_CUDA(cudaMemcpyAsync(host, dev, sizeof(double) * count, cudaMemcpyDeviceToHost, c.stream));
    NEKB_CUDA(cudaStreamSynchronize(c.stream));
    return host[0];
}

inline double gm_glsc3(const double *a, const double *b, const double *m, int64_t n)
{
    Ctx &c = ctx();
    GmresState &G = gmres_state();
    G.scal.ensure(64);
    gm_sum3_kernel<<<cg_grid(n), 256, 0, c.stream>>>(a, b, m, n, G.scal.p, c.partials.p, &c.sc.p->counter[0]);
    NEKB_LAUNCHED();
    double h = 0.0;
    return gm_reduce_to_host(G.scal.p, 1, &h);
}

// ortho (core/navier1.f:223-257): remove the plain mean over all local entries when the pressure has a null space
inline void gm_ortho(double *x, int64_t n)
{
    GmresState &G = gmres_state();
    if (!G.ifvcor) return;
    NEKB_REQUIRE(G.ntotg > 0, "hmh_gmres: ifvcor set but the global point count was not registered");
    const double sum = gm_glsc3(x, nullptr, nullptr, n);
    gm_cadd_kernel<<<cg_grid(n), 256, 0, ctx().stream>>>(x, -sum / G.ntotg, n);
    NEKB_LAUNCHED();
}

// w = pmask * dssum(A z)   (gmres.f:284-302)
inline void gm_ax(double *w, const double *z, const double *h1, const double *h2, const double *pmask, int nel, int gs_handle)
{
    launch_ax(z, w, h1, h2, nel, nullptr);
    gs_op(gs_handle, w, 1, pmask);
}

// chktcg1 (core/hmholtz.f:527-609), imesh = 1, double precision (EPS = 1e-13), EIGAA = 0 => ACONDNO = 10
inline double chktcg1_dev(double tol, const double *res, const double *h1, const double *h2, const double *mask, const double *mult,
                          const double *binv, int nel, double vol)
{
    Ctx &c = ctx();
    GmresState &G = gmres_state();
    const int64_t n = (int64_t)nel * c.nxyz;
    const double eps = 1.0e-13, acondno = 10.0;
    G.w.ensure((size_t)n), G.r.ensure((size_t)n);
    gm_cmult2_kernel<<<cg_grid(n), 256, 0, c.stream>>>(G.w.p, res, 1.0, n);  // w2 = binv*res below via sum3
    NEKB_LAUNCHED();
    col2_kernel<<<cg_grid(n), 256, 0, c.stream>>>(G.w.p, binv, n);
    NEKB_LAUNCHED();
    const double rinit = sqrt(gm_glsc3(G.w.p, res, mult, n) / vol);
    const double rmin = eps * rinit;
    if (tol < rmin) tol = rmin;
    const double bcneu1 = gm_glsc3(mask, mult, nullptr, n), bcneu2 = gm_glsc3(mult, nullptr, nullptr, n);
    const double bctest = fabs(bcneu1 - bcneu2);
    fill_kernel<<<cg_grid(n), 256, 0, c.stream>>>(G.w.p, 1.0, n);
    NEKB_LAUNCHED();
    launch_ax(G.w.p, G.r.p, h1, h2, nel, nullptr);
    const double bcrob = sqrt(gm_glsc3(G.r.p, G.r.p, c.bm1.p, n) / vol);
    if (bctest < 0.1 && bcrob < eps * acondno) {
        const double tolmin = rinit * eps * 10.0;
        if (tol < tolmin) tol = tolmin;
    }
    return tol;
}

// Returns the iteration count.  tol > 0: absolute tolerance on rnorm = |gamma_(j+1)| / sqrt(vol) (tolpss);
// tol < 0: |tol| * div0 (param(21) < 0, gmres.f:376).  res is overwritten with the solution (:531).
// hist_host (may be NULL, maxit+1 doubles): rnorm per iteration; div0_out (may be NULL).
inline int hmh_gmres_run(double *res, const double *h1, const double *h2, const double *wt, const double *pmask, int nel,
                         int gs_handle, double vol, double tol, int maxit, double *hist_host, double *div0_out)
{
    Ctx &c = ctx();
    GmresState &G = gmres_state();
    cudaStream_t s = c.stream;
    const int64_t n = (int64_t)nel * c.nxyz;
    const int m = G.m, grid = cg_grid(n);
    if ((int)G.V.size() != m + 1) G.V.resize(m + 1), G.Z.resize(m);
    G.w.ensure((size_t)n), G.r.ensure((size_t)n), G.x.ensure((size_t)n), G.scal.ensure(64);
    c.partials.ensure((size_t)(GM_NV > 4 ? GM_NV : 4) * CG_PART_STRIDE);
    auto Vj = [&](int j) -> double * {
        G.V[j].ensure((size_t)n);
        return G.V[j].p;
    };
    auto Zj = [&](int j) -> double * {
        G.Z[j].ensure((size_t)n);
        return G.Z[j].p;
    };
    const double norm_fac = 1.0 / sqrt(vol);
    std::vector<double> H((size_t)(m + 1) * m, 0.0), cg(m, 0.0), sg(m, 0.0), gam(m + 1, 0.0), cvec(m, 0.0), hcol(m + 1);
    auto Hm = [&](int i, int j) -> double & { return H[(size_t)i * m + j]; };
    NEKB_CUDA(cudaMemsetAsync(G.x.p, 0, sizeof(double) * (size_t)n, s));
    int iter = 0, j = 0;
    bool conv = false;
    double div0 = 0.0, tolpss = tol, rnorm = 0.0;
    unsigned *counter = &c.sc.p->counter[0];
    c.last_hist.assign((size_t)(maxit > 0 ? maxit : 0) + 1, 0.0);   // rnorm per iteration, for nekb_last_history
    c.last_hist_rows = 0, c.last_hist_cols = 1;
    while (!conv) {
        if (iter == 0) {
            NEKB_CUDA(cudaMemcpyAsync(G.r.p, res, sizeof(double) * (size_t)n, cudaMemcpyDeviceToDevice, s));  // :352
        } else {  // :355-360
            gm_ax(G.w.p, G.x.p, h1, h2, pmask, nel, gs_handle);
            gm_sub3_kernel<<<grid, 256, 0, s>>>(G.r.p, res, G.w.p, n);
            NEKB_LAUNCHED();
        }
        gam[0] = sqrt(gm_glsc3(G.r.p, G.r.p, wt, n));  // :363
        if (iter == 0) {
            div0 = gam[0] * norm_fac;
            if (tol < 0) tolpss = fabs(tol) * div0;
        }
        rnorm = 0.0;
        if (gam[0] == 0.0) break;  // :371 lucky convergence
        gm_cmult2_kernel<<<grid, 256, 0, s>>>(Vj(0), G.r.p, 1.0 / gam[0], n);
        NEKB_LAUNCHED();
        bool restart = false;
        for (j = 0; j < m; j++) {
            iter++;
            // w = v_j (mu = 1) ; z_j = M^-1 w ; the preconditioner masks its input in place, so hand it a copy
            NEKB_CUDA(cudaMemcpyAsync(G.w.p, Vj(j), sizeof(double) * (size_t)n, cudaMemcpyDeviceToDevice, s));
            h1mg_solve_dev(Zj(j), G.w.p);
            gm_ortho(Zj(j), n);
            gm_ax(G.w.p, Zj(j), h1, h2, pmask, nel, gs_handle);  // :410
            // h(i,j) = (w, v_i), i <= j : one reduction per column (:417-422)
            for (int i0 = 0; i0 <= j; i0 += GM_NV) {
                const int nv = (j + 1 - i0) < GM_NV ? (j + 1 - i0) : GM_NV;
                GmPtrs P;
                for (int q = 0; q < GM_NV; q++) P.v[q] = Vj(i0 + (q < nv ? q : 0)), P.h[q] = 0.0;
                gm_dots_kernel<<<grid, 256, 0, s>>>(G.w.p, wt, P, nv, n, G.scal.p + i0, c.partials.p, counter);
                NEKB_LAUNCHED();
            }
            gm_reduce_to_host(G.scal.p, j + 1, hcol.data());
            for (int i = 0; i <= j; i++) Hm(i, j) = hcol[i];
            double alpha2 = 0.0;
            for (int i0 = 0; i0 <= j; i0 += GM_NV) {  // :424-426 (+ :448 fused into the last pass)
                const int nv = (j + 1 - i0) < GM_NV ? (j + 1 - i0) : GM_NV;
                const bool last = i0 + GM_NV > j;
                GmPtrs P;
                for (int q = 0; q < GM_NV; q++) P.v[q] = Vj(i0 + (q < nv ? q : 0)), P.h[q] = q < nv ? hcol[i0 + q] : 0.0;
                gm_project_kernel<<<grid, 256, 0, s>>>(G.w.p, wt, P, nv, n, last ? 1 : 0, G.scal.p + 40, c.partials.p, counter);
                NEKB_LAUNCHED();
            }
            gm_reduce_to_host(G.scal.p + 40, 1, &alpha2);
            for (int i = 0; i < j; i++) {  // :441-447 Givens rotations of the new column
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
            const double l = sqrt(Hm(j, j) * Hm(j, j) + alpha * alpha);
            const double t = 1.0 / l;
            cg[j] = Hm(j, j) * t;
            sg[j] = alpha * t;
            Hm(j, j) = l;
            gam[j + 1] = -sg[j] * gam[j];
            gam[j] = cg[j] * gam[j];
            rnorm = fabs(gam[j + 1]) * norm_fac;
            if (hist_host) hist_host[iter - 1] = rnorm;
            if ((int)c.last_hist.size() >= iter) c.last_hist[iter - 1] = rnorm, c.last_hist_rows = iter;
            if (iter + 1 > maxit || rnorm < tolpss) {  // :466-467
                conv = true;
                break;
            }
            if (j == m - 1) {
                restart = true;
                break;
            }
            gm_cmult2_kernel<<<grid, 256, 0, s>>>(Vj(j + 1), G.w.p, 1.0 / alpha, n);  // :474
            NEKB_LAUNCHED();
        }
        (void)restart;
        const int kk = j + 1 > m ? m : j + 1;
        for (int k = kk - 1; k >= 0; k--) {  // :481-487 back substitution
            double t = gam[k];
            for (int i = kk - 1; i > k; i--) t = t - Hm(k, i) * cvec[i];
            cvec[k] = t / Hm(k, k);
        }
        for (int i0 = 0; i0 < kk; i0 += GM_NV) {  // :489-491
            const int nv = (kk - i0) < GM_NV ? (kk - i0) : GM_NV;
            GmPtrs P;
            for (int q = 0; q < GM_NV; q++) P.v[q] = Zj(i0 + (q < nv ? q : 0)), P.h[q] = q < nv ? cvec[i0 + q] : 0.0;
            gm_combine_kernel<<<grid, 256, 0, s>>>(G.x.p, P, nv, n);
            NEKB_LAUNCHED();
        }
    }
    NEKB_CUDA(cudaMemcpyAsync(res, G.x.p, sizeof(double) * (size_t)n, cudaMemcpyDeviceToDevice, s));  // :531
    gm_ortho(res, n);
    NEKB_CUDA(cudaStreamSynchronize(s));
    if (div0_out) *div0_out = div0;
    return iter;
}

// ---------------------------------------------------------------------------------------------- cggo('PRES'), param(42) = 1
// The two pieces the plain-PCG pressure solve adds to the Schwarz branch of cggo (core/hmholtz.f:737-748; cg.cuh cggo_run):
//
// crs_solve_h1 (core/navier8.f:1490-1535): w = J crs_solve( J^T (vmult * r) ) with the bilinear (vertex) basis
// h1_basis(i, 0..1) = (1 -+ z_i) / 2 (set_h1_basis_bilin :1537-1550; map_f_to_c_h1_bilin :1605-1646 and map_c_to_f_h1_bilin
// :1555-1603 are the same r, s, t contraction sequence as the multigrid transfers, so mg_tensor3 serves) and the coarse solver
// of the registered pressure multigrid (h1mg_setup built it from the same set_up_h1_crs data).  vmult is COMMON /SOLN/ vmult
// when registered, else the mult the caller handed to cggo.
struct PresCoarse {
    DevBuf<double> J;        // [lx1][2]
    DevBuf<double> vc, uc;   // [nel][8]
    int lx1 = 0;
};
inline PresCoarse &pres_coarse()
{
    static PresCoarse p;
    return p;
}
inline void cggo_pres_coarse(double *w, const double *r, const double *mult, int nel)
{
    Ctx &c = ctx();
    H1mg &M = h1mg();
    NEKB_REQUIRE(M.ready && !M.pnpn2 && M.nel == nel,
                 "cggo('PRES'), param(42) = 1: crs_solve_h1 needs the pressure coarse solver (nekb_h1mg_setup) on the same elements");
    PresCoarse &P = pres_coarse();
    const int nx = c.nx;
    if (P.lx1 != nx) {
        std::vector<double> J((size_t)nx * 2);
        for (int i = 0; i < nx; i++) J[(size_t)i * 2] = 0.5 * (1.0 - c.z_host[i]), J[(size_t)i * 2 + 1] = 0.5 * (1.0 + c.z_host[i]);
        P.J.upload(J.data(), J.size(), c.stream);
        P.lx1 = nx;
    }
    P.vc.ensure((size_t)nel * 8), P.uc.ensure((size_t)nel * 8);
    const int64_t n = (int64_t)nel * c.nxyz;
    const double *vm = c.vmult.n >= (size_t)n ? c.vmult.p : mult;
    mg_tensor3(P.vc.p, r, vm, P.J.p, 2, nx, true, false, nel);    // col3(uf,vf,vmult) ; map_f_to_c_h1_bilin
    crs_solve_dev(M, P.uc.p, P.vc.p);                             // crs_solve(xxth(ifield),uc,vc)
    mg_tensor3(w, P.uc.p, nullptr, P.J.p, nx, 2, false, false, nel);   // map_c_to_f_h1_bilin
}
// ortho (core/navier1.f:223-256) on the velocity-mesh pressure: the mean over all nelgv * lx1^3 entries goes when ifvcor is set
inline void cggo_pres_ortho(double *z, int64_t n) { gm_ortho(z, n); }

}  // namespace nekb
