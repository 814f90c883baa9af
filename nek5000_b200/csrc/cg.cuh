// cg.cuh -- device-resident conjugate-gradient drivers.
//
// cggos : examples/bp5/bp5.usr:797-899 (identity preconditioner, fixed iteration count when tol<0)
// cggo  : core/hmholtz.f:611-846, Jacobi branch (kfldfdm<0), incl. the exit rule of :778
//
// All CG scalars live on the device (struct CgScalars); the reductions are two-level with a fixed
// combination tree (common.cuh grid_reduce), the closing kernel's last block advances the iteration
// counter, so one iteration is a fixed sequence of launches with constant arguments: no host
// synchronisation inside the loop, and the sequence can be replayed from a CUDA graph.
#pragma once
#include "ax.cuh"
#include "gs.cuh"

namespace nekb {

inline void comm_allreduce_sum(double *dev, int count);  // comm.cuh (no-op on one rank)
inline void comm_allreduce_max(double *dev, int count);
// fdm_h1.cuh: the Schwarz branch of cggo (kfldfdm >= 0, core/hmholtz.f:731-746)
inline int fdm_h1_kfldfdm();
inline void set_fdm_prec_h1b_dev(double *d, const double *h1, const double *h2, int nel);
inline void fdm_h1_apply(double *z, const double *r, const double *d, const double *mask, int nel, int gs_handle);
inline void cggo_pres_coarse(double *w, const double *r, const double *mult, int nel);  // gmres.cuh: crs_solve_h1 (navier8.f:1490-1535)
inline void cggo_pres_ortho(double *z, int64_t n);                                      // gmres.cuh: ortho (navier1.f:223-256)

constexpr int CG_THREADS = 256;
constexpr int CG_PART_STRIDE = 1024;  // partial-sum region per kernel kind

inline int cg_grid(int64_t n)
{
    int64_t b = (n / 2 + CG_THREADS - 1) / CG_THREADS;
    int64_t cap = (int64_t)ctx().num_sms * 8;
    if (cap > CG_PART_STRIDE) cap = CG_PART_STRIDE - (CG_PART_STRIDE % ctx().num_sms);
    if (b > cap) b = cap;
    if (b < 1) b = 1;
    return (int)b;
}

// ---------------------------------------------------------------------------------------------- cggos
// bp5.usr:835-845: u=0, r=rhs, p=dpc*r (dpc=1), rpp1 = sum rmult*p*r
__global__ void __launch_bounds__(CG_THREADS)
    cggos_init_kernel(double *__restrict__ u, double *__restrict__ r, double *__restrict__ p,
                      const double *__restrict__ rhs, const double *__restrict__ mult, int64_t n, CgScalars *sc,
                      double *partials)
{
    __shared__ double red[33];
    double s = 0.0;
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < n; t += (int64_t)gridDim.x * blockDim.x) {
        const double rv = rhs[t];
        u[t] = 0.0;
        r[t] = rv;
        p[t] = rv;
        s = fma(mult[t] * rv, rv, s);
    }
    double b = block_reduce(s, red);
    grid_reduce(b, partials, &sc->counter[1], red, [=](double tot) {
        sc->rtz1 = tot;
        sc->it = 0;
        sc->done = 0;
    });
}

// bp5.usr:853-866: alpha = rpp1/pap ; u += alpha p ; r -= alpha (mask*ap) ; rz = sum rmult*r*r ;
// optional err = max|u-x1|.  (xmask1 of :850 is applied here instead of in a separate sweep.)
template <bool ERR>
__global__ void __launch_bounds__(CG_THREADS)
    cggos_update_kernel(double *__restrict__ u, double *__restrict__ r, const double *__restrict__ p,
                        const double *__restrict__ ap, const double *__restrict__ mask,
                        const double *__restrict__ mult, const double *__restrict__ x1, int64_t n, CgScalars *sc,
                        double *partials, double *hist)
{
    __shared__ double red[33];
    const double pap = sc->work[0];
    const double alpha = sc->rtz1 / pap;
    double s = 0.0, emax = 0.0;
    const int64_t n2 = n >> 1;
    double2 *u2 = reinterpret_cast<double2 *>(u), *r2 = reinterpret_cast<double2 *>(r);
    const double2 *p2 = reinterpret_cast<const double2 *>(p), *ap2 = reinterpret_cast<const double2 *>(ap),
                  *mk2 = reinterpret_cast<const double2 *>(mask), *mu2 = reinterpret_cast<const double2 *>(mult),
                  *x2 = reinterpret_cast<const double2 *>(x1);
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < n2; t += (int64_t)gridDim.x * blockDim.x) {
        double2 uv = u2[t], rv = r2[t];
        const double2 pv = p2[t], av = ap2[t], mk = mk2[t], mu = mu2[t];
        uv.x = fma(alpha, pv.x, uv.x);
        uv.y = fma(alpha, pv.y, uv.y);
        rv.x = fma(-alpha, av.x * mk.x, rv.x);
        rv.y = fma(-alpha, av.y * mk.y, rv.y);
        u2[t] = uv;
        r2[t] = rv;
        s = fma(mu.x * rv.x, rv.x, s);
        s = fma(mu.y * rv.y, rv.y, s);
        if (ERR) {
            const double2 xv = x2[t];
            emax = fmax(emax, fmax(fabs(uv.x - xv.x), fabs(uv.y - xv.y)));
        }
    }
    if ((n & 1) && blockIdx.x == 0 && threadIdx.x == 0) {  // odd tail (never for lx1 even)
        const int64_t t = n - 1;
        u[t] = fma(alpha, p[t], u[t]);
        r[t] = fma(-alpha, ap[t] * mask[t], r[t]);
        s = fma(mult[t] * r[t], r[t], s);
        if (ERR) emax = fmax(emax, fabs(u[t] - x1[t]));
    }
    double b = block_reduce(s, red);
    if (ERR) {
        double bm = block_reduce<true>(emax, red);
        grid_reduce<true>(bm, partials + CG_PART_STRIDE, &sc->counter[3], red, [=](double tot) { sc->work[1] = tot; });
    }
    grid_reduce(b, partials, &sc->counter[2], red, [=](double tot) { sc->rtz2 = tot; });  // new (r,z)
}

// bp5.usr:879-884: beta = rpp1/rpp2 ; p = z + beta p (z = r).  The last block to finish rotates the
// scalars and advances the iteration counter.
__global__ void __launch_bounds__(CG_THREADS)
    cggos_pupdate_kernel(double *__restrict__ p, const double *__restrict__ r, int64_t n, CgScalars *sc, double *hist)
{
    const double rz_old = sc->rtz1, rz_new = sc->rtz2;
    const double beta = rz_new / rz_old;
    const int64_t n2 = n >> 1;
    double2 *p2 = reinterpret_cast<double2 *>(p);
    const double2 *r2 = reinterpret_cast<const double2 *>(r);
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < n2; t += (int64_t)gridDim.x * blockDim.x) {
        double2 pv = p2[t];
        const double2 rv = r2[t];
        pv.x = fma(beta, pv.x, rv.x);
        pv.y = fma(beta, pv.y, rv.y);
        p2[t] = pv;
    }
    if ((n & 1) && blockIdx.x == 0 && threadIdx.x == 0) p[n - 1] = fma(beta, p[n - 1], r[n - 1]);
    __shared__ int s_last;
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        unsigned t = atomicInc(&sc->counter[1], gridDim.x - 1);
        s_last = (t == gridDim.x - 1);
    }
    __syncthreads();
    if (s_last && threadIdx.x == 0) {
        hist[3 * sc->it + 0] = sc->work[0];  // pap (after the all-reduce)
        hist[3 * sc->it + 1] = rz_new;
        sc->rtz1 = rz_new;
        sc->it = sc->it + 1;
    }
}

struct CggosArgs {
    double *u;
    const double *rhs, *x1, *mult, *mask;
    int gs_handle;
    int nel;
};
inline int cggos_run_fused(const CggosArgs &a, int maxit, double *hist_host, bool &used);  // below

// Runs the BP5 solver; returns iterations performed.  hist_host: 3 doubles per iteration
// (pap, rtz_new, max|u-x1| or 0).
inline int cggos_run(const CggosArgs &a, double tol, int maxit, double *hist_host)
{
    Ctx &c = ctx();
    if (!(tol > 0.0)) {  // fixed iteration count (bp5.par: tol < 0): no per-iteration host decision needed
        bool used = false;
        const int it = cggos_run_fused(a, maxit, hist_host, used);
        if (used) return it;
    }
    const int64_t n = (int64_t)a.nel * c.nxyz;
    cudaStream_t s = c.stream;
    DevBuf<double> &r = c.work[0], &p = c.work[1], &ap = c.work[2];
    r.ensure(n);
    p.ensure(n);
    ap.ensure(n);
    c.hist.ensure((size_t)3 * (maxit + 1));
    NEKB_CUDA(cudaMemsetAsync(c.hist.p, 0, sizeof(double) * 3 * (maxit + 1), s));
    c.partials.ensure(4 * CG_PART_STRIDE);
    CgScalars *sc = c.sc.p;
    const int grid = cg_grid(n);
    const bool err = tol > 0.0;
    GsMap &h = gs_get(a.gs_handle);
    NEKB_REQUIRE(h.n == n, "cggos: gs handle was set up for a different vector length");

    cggos_init_kernel<<<grid, CG_THREADS, 0, s>>>(a.u, r.p, p.p, a.rhs, a.mult, n, sc, c.partials.p + 1 * CG_PART_STRIDE);
    NEKB_LAUNCHED();
    comm_allreduce_sum(&sc->rtz1, 1);

    int iters = 0;
    bool broke = false;
    double last_pap = 0.0, last_rz = 0.0;
    for (int iter = 1; iter <= maxit; iter++) {
        prof_begin(PROF_AX);
        launch_ax(p.p, ap.p, nullptr, nullptr, a.nel, &sc->work[0]);
        prof_end(PROF_AX);
        comm_allreduce_sum(&sc->work[0], 1);
        prof_begin(PROF_GS);
        gs_op(a.gs_handle, ap.p, 1, nullptr);
        prof_end(PROF_GS);
        prof_begin(PROF_UPDATE);
        if (err)
            cggos_update_kernel<true><<<grid, CG_THREADS, 0, s>>>(a.u, r.p, p.p, ap.p, a.mask, a.mult, a.x1, n, sc,
                                                                  c.partials.p + 2 * CG_PART_STRIDE, c.hist.p);
        else
            cggos_update_kernel<false><<<grid, CG_THREADS, 0, s>>>(a.u, r.p, p.p, ap.p, a.mask, a.mult, a.x1, n, sc,
                                                                   c.partials.p + 2 * CG_PART_STRIDE, c.hist.p);
        NEKB_LAUNCHED();
        prof_end(PROF_UPDATE);
        comm_allreduce_sum(&sc->rtz2, 1);
        iters = iter;
        if (err) {  // bp5.usr:869-874: exit on enorm < tol needs the host in the loop
            comm_allreduce_max(&sc->work[1], 1);
            double e = 0.0;
            NEKB_CUDA(cudaMemcpyAsync(&e, &sc->work[1], sizeof(double), cudaMemcpyDeviceToHost, s));
            NEKB_CUDA(cudaStreamSynchronize(s));
            if (hist_host) hist_host[3 * (iter - 1) + 2] = e;
            if (e < tol) {  // the closing kernel is skipped: record this iteration's scalars here
                CgScalars hs;
                NEKB_CUDA(cudaMemcpyAsync(&hs, sc, sizeof(CgScalars), cudaMemcpyDeviceToHost, s));
                NEKB_CUDA(cudaStreamSynchronize(s));
                last_pap = hs.work[0];
                last_rz = hs.rtz2;
                broke = true;
                break;
            }
        }
        prof_begin(PROF_PUPDATE);
        cggos_pupdate_kernel<<<grid, CG_THREADS, 0, s>>>(p.p, r.p, n, sc, c.hist.p);
        NEKB_LAUNCHED();
        prof_end(PROF_PUPDATE);
    }
    if (prof().on) {
        NEKB_CUDA(cudaStreamSynchronize(s));
        prof_collect();
    }
    if (hist_host) {
        std::vector<double> hh((size_t)3 * (maxit + 1));
        NEKB_CUDA(cudaMemcpyAsync(hh.data(), c.hist.p, sizeof(double) * hh.size(), cudaMemcpyDeviceToHost, s));
        NEKB_CUDA(cudaStreamSynchronize(s));
        for (int i = 0; i < iters; i++) {
            hist_host[3 * i + 0] = hh[3 * i + 0];
            hist_host[3 * i + 1] = hh[3 * i + 1];
            if (!err) hist_host[3 * i + 2] = 0.0;
        }
        if (broke) {
            hist_host[3 * (iters - 1) + 0] = last_pap;
            hist_host[3 * (iters - 1) + 1] = last_rz;
        }
    }
    return iters;
}

// ---------------------------------------------------------------------------------------------- cggos, fused path
// Iteration = ax_cg_kernel (u, p update + Ax + pap)  ->  gs  ->  cggos_update2_kernel (r update + weighted dot).
// The multiplicity weight and the Dirichlet mask are folded into one byte per node: weights are 1/m with a small
// integer m (vmult = 1/dssum(1), core/connect1.f:124-135), the mask is 0/1 (bp5.usr:142-153), so the byte
// (m | mask==0 ? 0x80 : 0) reproduces both exactly; inputs that do not fit this form use the unfused path.
__global__ void __launch_bounds__(CG_THREADS)
    encode_weights_kernel(unsigned char *__restrict__ code, const double *__restrict__ mult,
                          const double *__restrict__ mask, int64_t n, int *bad)
{
    int b = 0;
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < n; t += (int64_t)gridDim.x * blockDim.x) {
        const double m = mult[t], k = mask[t];
        const double c = rint(1.0 / m);
        const bool okm = c >= 1.0 && c <= 127.0 && 1.0 / c == m;
        const bool okk = k == 0.0 || k == 1.0;
        if (!(okm && okk)) b = 1;
        code[t] = (unsigned char)((okm ? (int)c : 0) | (k == 0.0 ? 0x80 : 0));
    }
    if (b) atomicOr(bad, 1);
}

// u = 0, r = rhs, (r,z) = sum mult r r  (bp5.usr:835-845; p = r is written by the first ax_cg_kernel)
__global__ void __launch_bounds__(CG_THREADS)
    cggos_init2_kernel(double *__restrict__ u, double *__restrict__ r, const double *__restrict__ rhs,
                       const double *__restrict__ mult, int64_t n, CgScalars *sc, double *partials)
{
    __shared__ double red[33];
    double s = 0.0;
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < n; t += (int64_t)gridDim.x * blockDim.x) {
        const double rv = rhs[t];
        u[t] = 0.0;
        r[t] = rv;
        s = fma(mult[t] * rv, rv, s);
    }
    double b = block_reduce(s, red);
    grid_reduce(b, partials, &sc->counter[1], red, [=](double tot) {
        sc->work[1] = tot;
        sc->rtz1 = tot;
        sc->alpha = 0.0;
        sc->it = 0;
        sc->done = 0;
    });
}

// bp5.usr:853-866 without the u update: alpha = rpp1/pap ; r -= alpha*(mask*ap) ; (r,z) = sum mult r r.
// The last block stores alpha, rotates the (r,z) scalars and advances the iteration counter.
__global__ void __launch_bounds__(CG_THREADS, 6)   // cg_grid launches 6 CTAs per SM: all of them must be resident at once
    cggos_update2_kernel(double *__restrict__ r, const double *__restrict__ ap, const unsigned char *__restrict__ code,
                         int64_t n, CgScalars *sc, double *partials)
{
    __shared__ double red[33];
    __shared__ double wtab[128];
    const double pap = sc->work[0], rz = sc->work[1];
    const double alpha = rz / pap;
    if (threadIdx.x < 128) wtab[threadIdx.x] = threadIdx.x ? 1.0 / (double)threadIdx.x : 0.0;
    __syncthreads();
    double s = 0.0;
    const int64_t n4 = n >> 2;
    double2 *r2 = reinterpret_cast<double2 *>(r);
    const double2 *a2 = reinterpret_cast<const double2 *>(ap);
    const uchar4 *c4 = reinterpret_cast<const uchar4 *>(code);
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < n4; t += (int64_t)gridDim.x * blockDim.x) {
        double2 ra = r2[2 * t], rb = r2[2 * t + 1];
        const double2 aa = a2[2 * t], ab = a2[2 * t + 1];
        const uchar4 c = c4[t];
        ra.x = fma(-alpha, (c.x & 0x80) ? 0.0 : aa.x, ra.x);
        ra.y = fma(-alpha, (c.y & 0x80) ? 0.0 : aa.y, ra.y);
        rb.x = fma(-alpha, (c.z & 0x80) ? 0.0 : ab.x, rb.x);
        rb.y = fma(-alpha, (c.w & 0x80) ? 0.0 : ab.y, rb.y);
        r2[2 * t] = ra;
        r2[2 * t + 1] = rb;
        s = fma(wtab[c.x & 0x7f] * ra.x, ra.x, s);
        s = fma(wtab[c.y & 0x7f] * ra.y, ra.y, s);
        s = fma(wtab[c.z & 0x7f] * rb.x, rb.x, s);
        s = fma(wtab[c.w & 0x7f] * rb.y, rb.y, s);
    }
    if (blockIdx.x == 0 && threadIdx.x == 0)
        for (int64_t t = n4 << 2; t < n; t++) {
            const unsigned char c = code[t];
            r[t] = fma(-alpha, (c & 0x80) ? 0.0 : ap[t], r[t]);
            s = fma(wtab[c & 0x7f] * r[t], r[t], s);
        }
    double b = block_reduce(s, red);
    grid_reduce(b, partials, &sc->counter[2], red, [=](double tot) {
        sc->rtz1 = rz;
        sc->work[1] = tot;
        sc->alpha = alpha;
        sc->it = sc->it + 1;
    });
}

// EXPERIMENT, off by default.  Mode 1 (every group gathered here) measured slower than the gs + update pair at E = 262,144
// (2.15 vs 0.61 + 0.66 ms per iteration, profiles/r1o_gs_fuse_experiment_v*.json; bit-identical results) -- the suspected
// cost is the edge / corner nodes (80 of 512 per element), each walking its group's CSR through three dependent, divergent
// loads.  Mode 2 therefore assembles those groups in place first (gs_local_kernel on goff3 / gidx3) and gathers only the
// face pairs here; it was written after the GPU budget was spent and has NOT been run yet.
// The same update with the direct-stiffness summation folded in (one rank, NEKB_GS_FUSE_UPDATE=1): `ap` holds the
// UN-assembled A p; every node gathers the members of its group on the fly (gs.cuh gs_gathered: the bits gs_op would
// produce), so the assembled vector is never written back or re-read -- the scattered partner reads are served by L2
// because the partners are streamed by neighbouring blocks of the same sweep.  Masked nodes skip the gather.
__global__ void __launch_bounds__(CG_THREADS)
    cggos_update2_gs_kernel(double *__restrict__ r, const double *__restrict__ ap, const unsigned char *__restrict__ code,
                            const int32_t *__restrict__ link, const int32_t *__restrict__ goff,
                            const int32_t *__restrict__ gidx, int64_t n, CgScalars *sc, double *partials)
{
    __shared__ double red[33];
    __shared__ double wtab[128];
    const double pap = sc->work[0], rz = sc->work[1];
    const double alpha = rz / pap;
    if (threadIdx.x < 128) wtab[threadIdx.x] = threadIdx.x ? 1.0 / (double)threadIdx.x : 0.0;
    __syncthreads();
    double s = 0.0;
    const int64_t n4 = n >> 2;
    double2 *r2 = reinterpret_cast<double2 *>(r);
    const double2 *a2 = reinterpret_cast<const double2 *>(ap);
    const uchar4 *c4 = reinterpret_cast<const uchar4 *>(code);
    const int4 *l4 = reinterpret_cast<const int4 *>(link);
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < n4; t += (int64_t)gridDim.x * blockDim.x) {
        double2 ra = r2[2 * t], rb = r2[2 * t + 1];
        const double2 aa = a2[2 * t], ab = a2[2 * t + 1];
        const uchar4 c = c4[t];
        const int4 l = l4[t];
        // the four partner loads are issued together (a node without a partner re-reads itself: an L1 hit), the rare
        // edge / corner groups take the loop afterwards
        const int64_t t4 = t << 2;
        const double q0 = ap[l.x >= 0 ? (int64_t)l.x : t4], q1 = ap[l.y >= 0 ? (int64_t)l.y : t4 + 1];
        const double q2 = ap[l.z >= 0 ? (int64_t)l.z : t4 + 2], q3 = ap[l.w >= 0 ? (int64_t)l.w : t4 + 3];
        double w0 = l.x >= 0 ? aa.x + q0 : aa.x, w1 = l.y >= 0 ? aa.y + q1 : aa.y;
        double w2 = l.z >= 0 ? ab.x + q2 : ab.x, w3 = l.w >= 0 ? ab.y + q3 : ab.y;
        if (l.x < -1) w0 = gs_gathered(ap, aa.x, l.x, goff, gidx);
        if (l.y < -1) w1 = gs_gathered(ap, aa.y, l.y, goff, gidx);
        if (l.z < -1) w2 = gs_gathered(ap, ab.x, l.z, goff, gidx);
        if (l.w < -1) w3 = gs_gathered(ap, ab.y, l.w, goff, gidx);
        if (c.x & 0x80) w0 = 0.0;
        if (c.y & 0x80) w1 = 0.0;
        if (c.z & 0x80) w2 = 0.0;
        if (c.w & 0x80) w3 = 0.0;
        ra.x = fma(-alpha, w0, ra.x);
        ra.y = fma(-alpha, w1, ra.y);
        rb.x = fma(-alpha, w2, rb.x);
        rb.y = fma(-alpha, w3, rb.y);
        r2[2 * t] = ra;
        r2[2 * t + 1] = rb;
        s = fma(wtab[c.x & 0x7f] * ra.x, ra.x, s);
        s = fma(wtab[c.y & 0x7f] * ra.y, ra.y, s);
        s = fma(wtab[c.z & 0x7f] * rb.x, rb.x, s);
        s = fma(wtab[c.w & 0x7f] * rb.y, rb.y, s);
    }
    if (blockIdx.x == 0 && threadIdx.x == 0)
        for (int64_t t = n4 << 2; t < n; t++) {
            const unsigned char c = code[t];
            const double w = (c & 0x80) ? 0.0 : gs_gathered(ap, ap[t], link[t], goff, gidx);
            r[t] = fma(-alpha, w, r[t]);
            s = fma(wtab[c & 0x7f] * r[t], r[t], s);
        }
    double b = block_reduce(s, red);
    grid_reduce(b, partials, &sc->counter[2], red, [=](double tot) {
        sc->rtz1 = rz;
        sc->work[1] = tot;
        sc->alpha = alpha;
        sc->it = sc->it + 1;
    });
}

// Default form of the update with the direct-stiffness summation folded in (structured gather, gs.cuh gs_ensure_struct):
// `ap` holds A p with only the "S" groups assembled in place; face-interior nodes add their partner through the 8-byte
// affine link of their face, edge / corner nodes read their group's value from gval, interior nodes are final as they are.
// Same members, same order as gs_local_kernel<1>, hence the same bits as the stock pair gs_op + cggos_update2_kernel.
// One thread = 4 consecutive nodes (half a row of an 8^3 tile: same j, k; i = 0..3 or 4..7).
__global__ void __launch_bounds__(CG_THREADS)
    cggos_update3_kernel(double *__restrict__ r, const double *__restrict__ ap, const unsigned char *__restrict__ code,
                         const FaceLink *__restrict__ ftab, const int32_t *__restrict__ etab, const double *__restrict__ gval,
                         int64_t n, CgScalars *sc, double *partials)
{
    __shared__ double red[33];
    __shared__ double wtab[128];
    const double pap = sc->work[0], rz = sc->work[1];
    const double alpha = rz / pap;
    if (threadIdx.x < 128) wtab[threadIdx.x] = threadIdx.x ? 1.0 / (double)threadIdx.x : 0.0;
    __syncthreads();
    double s = 0.0;
    const int64_t n4 = n >> 2;   // n is a multiple of 512
    double2 *r2 = reinterpret_cast<double2 *>(r);
    const double2 *a2 = reinterpret_cast<const double2 *>(ap);
    const uchar4 *c4 = reinterpret_cast<const uchar4 *>(code);
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < n4; t += (int64_t)gridDim.x * blockDim.x) {
        double2 ra = r2[2 * t], rb = r2[2 * t + 1];
        const double2 aa = a2[2 * t], ab = a2[2 * t + 1];
        const uchar4 c = c4[t];
        double w[4] = {aa.x, aa.y, ab.x, ab.y};
        const int l4 = (int)(t & 127);                 // position of the quad in its tile
        const int64_t el = t >> 7;
        const int i0 = (l4 & 1) << 2, j = (l4 >> 1) & 7, k = l4 >> 4;
        const int bj = (j == 0 || j == 7), bk = (k == 0 || k == 7);
        if (bj + bk == 0) {
            // only the row end (i = 0 or 7) is on the surface: a node of an x-face
            const int ie = i0 ? 3 : 0;                 // index in the quad of the surface node
            const FaceLink L = ftab[el * 6 + (i0 ? 1 : 0)];
            if (L.base >= 0 && !(((const unsigned char *)&c)[ie] & 0x80)) w[ie] += ap[(int64_t)L.base + j * L.sa + k * L.sb];
        } else if (bj + bk == 1) {
            // a y- or z-face row: the three inner nodes of the quad are face-interior, the row end is an edge node
            const int f = bj ? 2 + (j == 7) : 4 + (k == 7);
            const int bco = bj ? k : j;                // second in-face coordinate (first is i)
            const FaceLink L = ftab[el * 6 + f];
            const int ie = i0 ? 3 : 0;
            if (L.base >= 0) {
                const int64_t pb = (int64_t)L.base + bco * L.sb;
#pragma unroll
                for (int q = 0; q < 4; q++)
                    if (q != ie) w[q] += ap[pb + (i0 + q) * L.sa];
            }
            const int g = etab[el * GS_ST_EDGE_SLOTS + gs_st_slot(i0 ? 7 : 0, j, k)];
            if (g >= 0) w[ie] = gval[g];
        } else {
            // an x-edge row (j and k on the surface): every node is an edge node, the row end a corner
#pragma unroll
            for (int q = 0; q < 4; q++) {
                const int g = etab[el * GS_ST_EDGE_SLOTS + gs_st_slot(i0 + q, j, k)];
                if (g >= 0) w[q] = gval[g];
            }
        }
        ra.x = fma(-alpha, (c.x & 0x80) ? 0.0 : w[0], ra.x);
        ra.y = fma(-alpha, (c.y & 0x80) ? 0.0 : w[1], ra.y);
        rb.x = fma(-alpha, (c.z & 0x80) ? 0.0 : w[2], rb.x);
        rb.y = fma(-alpha, (c.w & 0x80) ? 0.0 : w[3], rb.y);
        r2[2 * t] = ra;
        r2[2 * t + 1] = rb;
        s = fma(wtab[c.x & 0x7f] * ra.x, ra.x, s);
        s = fma(wtab[c.y & 0x7f] * ra.y, ra.y, s);
        s = fma(wtab[c.z & 0x7f] * rb.x, rb.x, s);
        s = fma(wtab[c.w & 0x7f] * rb.y, rb.y, s);
    }
    double b = block_reduce(s, red);
    grid_reduce(b, partials, &sc->counter[2], red, [=](double tot) {
        sc->rtz1 = rz;
        sc->work[1] = tot;
        sc->alpha = alpha;
        sc->it = sc->it + 1;
    });
}

// Branch-free form of cggos_update3_kernel: every thread runs the SAME instruction stream over the four nodes of its quad --
// two predicated table loads per node (face link / edge slot), then two predicated value loads (partner / gval) -- so the
// 2 x 8 irregular loads of a thread are issued in two independent batches instead of inside divergent, serialised branches,
// and no shared memory or barrier is involved.  Bits as the other forms.
__global__ void __launch_bounds__(CG_THREADS, 4)
    cggos_update5_kernel(double *__restrict__ r, const double *__restrict__ ap, const unsigned char *__restrict__ code,
                         const FaceLink *__restrict__ ftab, const int32_t *__restrict__ etab, const double *__restrict__ gval,
                         int64_t n, CgScalars *sc, double *partials)
{
    __shared__ double red[33];
    __shared__ double wtab[128];
    const double pap = sc->work[0], rz = sc->work[1];
    const double alpha = rz / pap;
    if (threadIdx.x < 128) wtab[threadIdx.x] = threadIdx.x ? 1.0 / (double)threadIdx.x : 0.0;
    __syncthreads();
    double s = 0.0;
    const int64_t n4 = n >> 2;
    double2 *r2 = reinterpret_cast<double2 *>(r);
    const double2 *a2 = reinterpret_cast<const double2 *>(ap);
    const uchar4 *c4 = reinterpret_cast<const uchar4 *>(code);
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < n4; t += (int64_t)gridDim.x * blockDim.x) {
        double2 ra = r2[2 * t], rb = r2[2 * t + 1];
        const double2 aa = a2[2 * t], ab = a2[2 * t + 1];
        const uchar4 c = c4[t];
        double w[4] = {aa.x, aa.y, ab.x, ab.y};
        const int l4 = (int)(t & 127);
        const int64_t el = t >> 7;
        const int i0 = (l4 & 1) << 2, j = (l4 >> 1) & 7, k = l4 >> 4;
        const int bj = (j == 0 || j == 7), bk = (k == 0 || k == 7);
        const int hj = j == 7, hk = k == 7;
        FaceLink L[4];
        int g[4];
        int fa[4], fb[4];
        // level 1: the tables
#pragma unroll
        for (int q = 0; q < 4; q++) {
            const int i = i0 + q;
            const int bi = (i == 0 || i == 7), hi = i == 7;
            const int nb = bi + bj + bk;
            const int f = bi ? hi : (bj ? 2 + hj : 4 + hk);
            fa[q] = bi ? j : i;                  // in-face coordinates, lower axis first: x-face (j,k), y-face (i,k), z-face (i,j)
            fb[q] = (bi || bj) ? k : j;
            L[q].base = -1, L[q].sa = 0, L[q].sb = 0;
            if (nb == 1) L[q] = ftab[el * 6 + f];
            g[q] = -1;
            if (nb >= 2) g[q] = etab[el * GS_ST_EDGE_SLOTS + gs_st_slot(i, j, k)];
        }
        // level 2: the values
        double pv[4], gv[4];
#pragma unroll
        for (int q = 0; q < 4; q++) {
            pv[q] = 0.0, gv[q] = 0.0;
            if (L[q].base >= 0) pv[q] = ap[(int64_t)L[q].base + fa[q] * L[q].sa + fb[q] * L[q].sb];
            if (g[q] >= 0) gv[q] = gval[g[q]];
        }
#pragma unroll
        for (int q = 0; q < 4; q++) {
            if (L[q].base >= 0) w[q] += pv[q];
            if (g[q] >= 0) w[q] = gv[q];
        }
        ra.x = fma(-alpha, (c.x & 0x80) ? 0.0 : w[0], ra.x);
        ra.y = fma(-alpha, (c.y & 0x80) ? 0.0 : w[1], ra.y);
        rb.x = fma(-alpha, (c.z & 0x80) ? 0.0 : w[2], rb.x);
        rb.y = fma(-alpha, (c.w & 0x80) ? 0.0 : w[3], rb.y);
        r2[2 * t] = ra;
        r2[2 * t + 1] = rb;
        s = fma(wtab[c.x & 0x7f] * ra.x, ra.x, s);
        s = fma(wtab[c.y & 0x7f] * ra.y, ra.y, s);
        s = fma(wtab[c.z & 0x7f] * rb.x, rb.x, s);
        s = fma(wtab[c.w & 0x7f] * rb.y, rb.y, s);
    }
    double b = block_reduce(s, red);
    grid_reduce(b, partials, &sc->counter[2], red, [=](double tot) {
        sc->rtz1 = rz;
        sc->work[1] = tot;
        sc->alpha = alpha;
        sc->it = sc->it + 1;
    });
}

// The same update, organised by element instead of by node: 128 threads own one 8^3 tile (one quad of nodes each).  The
// irregular part -- 216 face partners and 80 edge / corner values per element -- is fetched by ALL threads in one uniform,
// fully independent sweep into shared memory (no divergent dependent loads in the streaming part), then every thread
// patches its quad from shared memory.  Bits as cggos_update3_kernel / the stock pair.
template <int EPB, int MINB>
__global__ void __launch_bounds__(128 * EPB, MINB)
    cggos_update4_kernel(double *__restrict__ r, const double *__restrict__ ap, const unsigned char *__restrict__ code,
                         const FaceLink *__restrict__ ftab, const int32_t *__restrict__ etab, const double *__restrict__ gval,
                         int nel, CgScalars *sc, double *partials)
{
    __shared__ double red[33];
    __shared__ double wtab[128];
    __shared__ double s_face[EPB][6 * 36];
    __shared__ double s_edge[EPB][GS_ST_EDGE_SLOTS];
    __shared__ int s_eg[EPB][GS_ST_EDGE_SLOTS];
    __shared__ int s_fok[EPB][6];
    const double pap = sc->work[0], rz = sc->work[1];
    const double alpha = rz / pap;
    if (threadIdx.x < 128) wtab[threadIdx.x] = threadIdx.x ? 1.0 / (double)threadIdx.x : 0.0;
    const int es = threadIdx.x >> 7, lt = threadIdx.x & 127;
    const int i0 = (lt & 1) << 2, j = (lt >> 1) & 7, k = lt >> 4;
    const int bj = (j == 0 || j == 7), bk = (k == 0 || k == 7);
    const int ie = i0 ? 3 : 0;
    double s = 0.0;
    double2 *r2 = reinterpret_cast<double2 *>(r);
    const double2 *a2 = reinterpret_cast<const double2 *>(ap);
    const uchar4 *c4 = reinterpret_cast<const uchar4 *>(code);
    for (int e0 = blockIdx.x * EPB; e0 < nel; e0 += gridDim.x * EPB) {
        const int e = e0 + es;
        const bool act = e < nel;
        const int64_t t = (int64_t)(act ? e : 0) * 128 + lt;
        double2 ra, rb, aa, ab;
        uchar4 c;
        if (act) {
            ra = r2[2 * t], rb = r2[2 * t + 1];
            aa = a2[2 * t], ab = a2[2 * t + 1];
            c = c4[t];
            // the irregular part, uniform over the threads: items 0..215 = (face, in-face node), 216..295 = edge / corner slots
#pragma unroll
            for (int it = lt; it < 216 + GS_ST_EDGE_SLOTS; it += 128) {
                if (it < 216) {
                    const int f = it / 36, q = it - 36 * f;
                    const int b = q / 6, a = q - 6 * b;
                    const FaceLink L = ftab[(int64_t)e * 6 + f];
                    double v = 0.0;
                    if (L.base >= 0) v = ap[(int64_t)L.base + (a + 1) * L.sa + (b + 1) * L.sb];
                    s_face[es][it] = v;
                    if (q == 0) s_fok[es][f] = L.base >= 0;
                } else {
                    const int sl = it - 216;
                    const int g = etab[(int64_t)e * GS_ST_EDGE_SLOTS + sl];
                    s_eg[es][sl] = g;
                    s_edge[es][sl] = g >= 0 ? gval[g] : 0.0;
                }
            }
        }
        __syncthreads();
        if (act) {
            double w[4] = {aa.x, aa.y, ab.x, ab.y};
            if (bj + bk == 0) {
                const int f = i0 ? 1 : 0;
                if (s_fok[es][f]) w[ie] += s_face[es][f * 36 + (j - 1) + 6 * (k - 1)];
            } else if (bj + bk == 1) {
                const int f = bj ? 2 + (j == 7) : 4 + (k == 7);
                const int bco = bj ? k : j;
                if (s_fok[es][f]) {
#pragma unroll
                    for (int q = 0; q < 4; q++)
                        if (q != ie) w[q] += s_face[es][f * 36 + (i0 + q - 1) + 6 * (bco - 1)];
                }
                const int sl = gs_st_slot(i0 ? 7 : 0, j, k);
                if (s_eg[es][sl] >= 0) w[ie] = s_edge[es][sl];
            } else {
#pragma unroll
                for (int q = 0; q < 4; q++) {
                    const int sl = gs_st_slot(i0 + q, j, k);
                    if (s_eg[es][sl] >= 0) w[q] = s_edge[es][sl];
                }
            }
            ra.x = fma(-alpha, (c.x & 0x80) ? 0.0 : w[0], ra.x);
            ra.y = fma(-alpha, (c.y & 0x80) ? 0.0 : w[1], ra.y);
            rb.x = fma(-alpha, (c.z & 0x80) ? 0.0 : w[2], rb.x);
            rb.y = fma(-alpha, (c.w & 0x80) ? 0.0 : w[3], rb.y);
            r2[2 * t] = ra;
            r2[2 * t + 1] = rb;
            s = fma(wtab[c.x & 0x7f] * ra.x, ra.x, s);
            s = fma(wtab[c.y & 0x7f] * ra.y, ra.y, s);
            s = fma(wtab[c.z & 0x7f] * rb.x, rb.x, s);
            s = fma(wtab[c.w & 0x7f] * rb.y, rb.y, s);
        }
        __syncthreads();   // the shared tables are rewritten by the next element
    }
    double b = block_reduce(s, red);
    grid_reduce(b, partials, &sc->counter[2], red, [=](double tot) {
        sc->rtz1 = rz;
        sc->work[1] = tot;
        sc->alpha = alpha;
        sc->it = sc->it + 1;
    });
}

// The same update as a streaming pipeline: ONE WARP owns an element; the element's r and A p tiles, its byte codes and its two
// gather tables (48 B of face links + 320 B of edge slots) arrive in a per-warp ring of shared-memory stages by TMA bulk copies
// (one mbarrier per stage), so the bytes in flight do not depend on how many threads are stalled behind a barrier -- ncu of
// cggos_update4_kernel showed DRAM 51-55 % active at near-minimal traffic: a latency chain (table -> partner -> barrier), not
// bandwidth.  The irregular part is software-pipelined: as soon as the tables of element n+1 are in shared memory, every lane
// issues the (<= 16, predicated) partner / gval loads of ITS OWN nodes of that element into registers, then computes element n
// with the values fetched one trip earlier.  A lane owns the two k-columns (i0, i0+1; j): one double2 per plane, 512 contiguous
// bytes per warp instruction for the shared-memory reads and the r stores.  No block barrier, no shared-memory staging of the
// gathered values.  Same members and member order per node as gs_local_kernel<1> (pairs: own + partner; edge / corner groups:
// gval), i.e. the bits of the other forms; (r,r) is summed in another grouping.
constexpr int UPD6_STAGE_BYTES = 2 * 4096 + 512 + 48 + 320 + 16;   // r, ap, code, ftab[6], etab[80]; padded to 128 B
static_assert(UPD6_STAGE_BYTES % 128 == 0, "stage alignment");

template <int WARPS, int STAGES>
__global__ void __launch_bounds__(32 * WARPS, 1)
    cggos_update6_kernel(double *__restrict__ r, const double *__restrict__ ap, const unsigned char *__restrict__ code,
                         const FaceLink *__restrict__ ftab, const int32_t *__restrict__ etab, const double *__restrict__ gval,
                         int nel, CgScalars *sc, double *partials)
{
    static_assert(STAGES >= 2, "the next element's tables must be resident while this one is computed");
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ double red[33];
    __shared__ double wtab[128];
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem_raw + (size_t)WARPS * STAGES * UPD6_STAGE_BYTES);
    const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
    unsigned char *wbase = smem_raw + (size_t)wid * STAGES * UPD6_STAGE_BYTES;
    uint64_t *full = bars + wid * STAGES;
    if (threadIdx.x == 0) {
        for (int q = 0; q < WARPS * STAGES; q++) mbar_init(&bars[q], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (threadIdx.x < 128) wtab[threadIdx.x] = threadIdx.x ? 1.0 / (double)threadIdx.x : 0.0;
    __syncthreads();
    const double pap = sc->work[0], rz = sc->work[1];
    const double alpha = rz / pap;

    // the lane's 16 nodes: (i0 + q, j, k), q = 0, 1, k = 0..7; what each of them gathers (fixed per lane)
    //   desc = -1: interior node; 0..5: face-interior node of face f, with in-face coordinates (fa, fb); 8 + slot: edge / corner
    const int i0 = (lane & 3) * 2, j = lane >> 2;
    int desc[16], fa[16], fb[16];
#pragma unroll
    for (int n = 0; n < 16; n++) {
        const int k = n >> 1, i = i0 + (n & 1);
        const int nb = gs_st_nb(i, j, k);
        desc[n] = -1, fa[n] = 0, fb[n] = 0;
        if (nb == 1) desc[n] = gs_st_face(i, j, k, fa[n], fb[n]);
        if (nb >= 2) desc[n] = 8 + gs_st_slot(i, j, k);
    }

    const int first = blockIdx.x * WARPS + wid, stride = gridDim.x * WARPS;
    auto issue = [&](int stage, int e) {
        unsigned char *st = wbase + (size_t)stage * UPD6_STAGE_BYTES;
        mbar_expect_tx(&full[stage], 2 * 4096 + 512 + 48 + 320);
        bulk_g2s(st, r + (size_t)e * 512, 4096, &full[stage]);
        bulk_g2s(st + 4096, ap + (size_t)e * 512, 4096, &full[stage]);
        bulk_g2s(st + 8192, code + (size_t)e * 512, 512, &full[stage]);
        bulk_g2s(st + 8704, ftab + (size_t)e * 6, 48, &full[stage]);
        bulk_g2s(st + 8752, etab + (size_t)e * GS_ST_EDGE_SLOTS, 320, &full[stage]);
    };
    if (lane == 0) {
#pragma unroll
        for (int s = 0; s < STAGES; s++) {
            const int e = first + s * stride;
            if (e < nel) issue(s, e);
        }
    }
    // gathers of the lane's own nodes of the element in `stage` (tables already in shared memory): values + validity mask
    auto gather = [&](int stage, double (&gv)[16]) -> unsigned {
        const unsigned char *st = wbase + (size_t)stage * UPD6_STAGE_BYTES;
        const FaceLink *sf = reinterpret_cast<const FaceLink *>(st + 8704);
        const int32_t *se = reinterpret_cast<const int32_t *>(st + 8752);
        unsigned vm = 0;
#pragma unroll
        for (int n = 0; n < 16; n++) {
            gv[n] = 0.0;
            if (desc[n] >= 8) {
                const int gI = se[desc[n] - 8];
                if (gI >= 0) {
                    gv[n] = gval[gI];
                    vm |= 1u << n;
                }
            } else if (desc[n] >= 0) {
                const FaceLink Lk = sf[desc[n]];
                if (Lk.base >= 0) {
                    gv[n] = ap[(int64_t)Lk.base + fa[n] * Lk.sa + fb[n] * Lk.sb];
                    vm |= 1u << n;
                }
            }
        }
        return vm;
    };

    double s = 0.0;
    double gcur[16];
    unsigned vcur = 0;
    if (first < nel) {
        mbar_wait(&full[0], 0);
        vcur = gather(0, gcur);
    }
    int it = 0;
    for (int e = first; e < nel; e += stride, it++) {
        const int stage = it % STAGES;
        // the next element's tables -> its gathers go out before this element is touched
        double gnext[16];
        unsigned vnext = 0;
        const int en1 = e + stride;
        if (en1 < nel) {
            const int st1 = (it + 1) % STAGES;
            mbar_wait(&full[st1], (uint32_t)((it + 1) / STAGES) & 1u);
            vnext = gather(st1, gnext);
        }
        const unsigned char *st = wbase + (size_t)stage * UPD6_STAGE_BYTES;
        const double2 *sr2 = reinterpret_cast<const double2 *>(st);
        const double2 *sa2 = reinterpret_cast<const double2 *>(st + 4096);
        const uchar2 *sc2 = reinterpret_cast<const uchar2 *>(st + 8192);
        double2 *r2 = reinterpret_cast<double2 *>(r + (size_t)e * 512);
#pragma unroll
        for (int k = 0; k < 8; k++) {
            const int q = k * 32 + lane;
            double2 rv = sr2[q];
            const double2 av = sa2[q];
            const uchar2 c = sc2[q];
            double w0 = av.x, w1 = av.y;
            if (vcur & (1u << (2 * k))) w0 = desc[2 * k] >= 8 ? gcur[2 * k] : w0 + gcur[2 * k];
            if (vcur & (1u << (2 * k + 1))) w1 = desc[2 * k + 1] >= 8 ? gcur[2 * k + 1] : w1 + gcur[2 * k + 1];
            rv.x = fma(-alpha, (c.x & 0x80) ? 0.0 : w0, rv.x);
            rv.y = fma(-alpha, (c.y & 0x80) ? 0.0 : w1, rv.y);
            r2[q] = rv;
            s = fma(wtab[c.x & 0x7f] * rv.x, rv.x, s);
            s = fma(wtab[c.y & 0x7f] * rv.y, rv.y, s);
        }
        __syncwarp();   // every lane is done with this stage
        if (lane == 0) {
            const int en = e + STAGES * stride;
            if (en < nel) issue(stage, en);   // the stage was only read: no generic-proxy writes to order
        }
#pragma unroll
        for (int n = 0; n < 16; n++) gcur[n] = gnext[n];
        vcur = vnext;
    }
    double b = block_reduce(s, red);
    grid_reduce(b, partials, &sc->counter[2], red, [=](double tot) {
        sc->rtz1 = rz;
        sc->work[1] = tot;
        sc->alpha = alpha;
        sc->it = sc->it + 1;
    });
}

template <int WARPS, int STAGES>
inline void launch_cggos_update6(double *r, const double *ap, const unsigned char *code, const GsMap &h, int nel, CgScalars *sc,
                                 double *partials)
{
    Ctx &c = ctx();
    constexpr size_t bytes = (size_t)WARPS * STAGES * UPD6_STAGE_BYTES + WARPS * STAGES * sizeof(uint64_t);
    static_assert(bytes + 2048 <= 227 * 1024, "shared memory of one SM");
    static bool configured = false;
    if (!configured) {
        NEKB_CUDA(cudaFuncSetAttribute(cggos_update6_kernel<WARPS, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
        configured = true;
    }
    cggos_update6_kernel<WARPS, STAGES><<<grid_for((nel + WARPS - 1) / WARPS, 1), 32 * WARPS, bytes, c.stream>>>(
        r, ap, code, h.ftab.p, h.etab.p, h.gval.p, nel, sc, partials);
}

// a += b (core/math.f add2)
__global__ void __launch_bounds__(CG_THREADS) add2_kernel(double *__restrict__ a, const double *__restrict__ b, int64_t n)
{
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < n; t += (int64_t)gridDim.x * blockDim.x) a[t] += b[t];
}

// u += alpha * p with the device-resident alpha (the u update of the final iteration)
__global__ void __launch_bounds__(CG_THREADS)
    axpy_alpha_kernel(double *__restrict__ u, const double *__restrict__ p, int64_t n, const CgScalars *sc)
{
    const double alpha = sc->alpha;
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < n; t += (int64_t)gridDim.x * blockDim.x)
        u[t] = fma(alpha, p[t], u[t]);
}

__global__ void cggos_hist_kernel(const CgScalars *sc, double *hist, int slot)
{
    hist[3 * slot + 0] = sc->work[0];
    hist[3 * slot + 1] = sc->work[1];
}

inline int axcg_variant()   // read per solve: the parity tests run the kept kernel forms side by side in one process
{
    const char *e = getenv("NEKB_AXCG_VARIANT");
    return e ? atoi(e) : 0;
}
inline int gs_fuse_update_enabled()  // read per solve, so one process can time both forms
{
    const char *e = getenv("NEKB_GS_FUSE_UPDATE");
    return e ? atoi(e) : 6;
}
inline int cg_fused_enabled()   // read per solve: bench.py times the kernel-per-statement path beside the fused one
{
    const char *e = getenv("NEKB_CG_FUSED");
    return e ? atoi(e) : 1;
}

inline int cggos_run_fused(const CggosArgs &a, int maxit, double *hist_host, bool &used)
{
    Ctx &c = ctx();
    used = false;
    if (c.nx != 8 || !cg_fused_enabled() || maxit < 1 || a.nel < 1) return 0;
    const int64_t n = (int64_t)a.nel * c.nxyz;
    cudaStream_t s = c.stream;
    const int grid = cg_grid(n);
    c.wcode.ensure((size_t)n);
    c.flags.ensure(4);
    NEKB_CUDA(cudaMemsetAsync(c.flags.p, 0, sizeof(int), s));
    encode_weights_kernel<<<grid, CG_THREADS, 0, s>>>(c.wcode.p, a.mult, a.mask, n, c.flags.p);
    NEKB_LAUNCHED();
    int bad = 0;
    NEKB_CUDA(cudaMemcpyAsync(&bad, c.flags.p, sizeof(int), cudaMemcpyDeviceToHost, s));
    NEKB_CUDA(cudaStreamSynchronize(s));
    if (bad) return 0;
    used = true;
    DevBuf<double> &r = c.work[0], &p = c.work[1], &ap = c.work[2];
    r.ensure(n), p.ensure(n), ap.ensure(n);
    c.hist.ensure((size_t)3 * (maxit + 1));
    if (hist_host) NEKB_CUDA(cudaMemsetAsync(c.hist.p, 0, sizeof(double) * 3 * (maxit + 1), s));
    c.partials.ensure(4 * CG_PART_STRIDE);
    CgScalars *sc = c.sc.p;
    GsMap &h = gs_get(a.gs_handle);
    NEKB_REQUIRE(h.n == n, "cggos: gs handle was set up for a different vector length");

    // gather form of dssum inside the update kernel: one rank only (no remote members to wait for)
    // (1 = every group gathered by the update kernel; 2 = pairs gathered, edge / corner groups assembled in place first)
    // 3 / 4 (default 4) = structured gather: face pairs through per-face affine links, edge / corner groups through gval,
    // groups with remote members in place + the inter-rank exchange (3: node-organised update kernel, 4: element-organised);
    // 0 = the stock pair gs_op + update2
    const int gmode = gs_fuse_update_enabled();
    const bool gather = (gmode == 1 || gmode == 2) && c.nranks == 1 && h.nshared == 0;
    if (gather) gs_ensure_link(h, gmode);
    const bool structured = (gmode >= 3 && gmode <= 6) && gs_ensure_struct(h);
    const bool affine = ax_affine_ensure();   // decided from the registered factors (all elements affine to 1e-13)
    cggos_init2_kernel<<<grid, CG_THREADS, 0, s>>>(a.u, r.p, a.rhs, a.mult, n, sc, c.partials.p + 1 * CG_PART_STRIDE);
    NEKB_LAUNCHED();
    comm_allreduce_sum(&sc->work[1], 1);
    const int axv = axcg_variant();
    for (int iter = 1; iter <= maxit; iter++) {
        prof_begin(PROF_AX);
        if (affine) {   // per-element constants instead of per-node factors (ax.cuh kernels v4 / v5); stages are 12 KB
            switch (axv) {
                case 1: launch_ax_cg_affine<8, 4, 4>(r.p, p.p, a.u, ap.p, a.nel, iter == 1, &sc->work[0]); break;
                case 2: launch_ax_cg_affine<8, 6, 3>(r.p, p.p, a.u, ap.p, a.nel, iter == 1, &sc->work[0]); break;
                case 3: launch_ax_cg_affine<8, 8, 2>(r.p, p.p, a.u, ap.p, a.nel, iter == 1, &sc->work[0]); break;
                case 4: launch_ax_cg_affine<8, 3, 4>(r.p, p.p, a.u, ap.p, a.nel, iter == 1, &sc->work[0]); break;
                case 5: launch_ax_cg_affine<8, 5, 3>(r.p, p.p, a.u, ap.p, a.nel, iter == 1, &sc->work[0]); break;
                case 6: launch_ax_cg_affine<8, 5, 2>(r.p, p.p, a.u, ap.p, a.nel, iter == 1, &sc->work[0]); break;
                // warp per element, in-plane contractions on DMMA (ax.cuh kernel v5): warps per SM x ring stages per warp
                case 11: launch_ax_cg_affine_mma<6, 3>(r.p, p.p, a.u, ap.p, a.nel, iter == 1, &sc->work[0]); break;
                case 14: launch_ax_cg_affine_mma<4, 4>(r.p, p.p, a.u, ap.p, a.nel, iter == 1, &sc->work[0]); break;
                case 7: launch_ax_cg_affine<8, 4, 3>(r.p, p.p, a.u, ap.p, a.nel, iter == 1, &sc->work[0]); break;   // round-2 default before v5
                default: launch_ax_cg_affine_mma<8, 2>(r.p, p.p, a.u, ap.p, a.nel, iter == 1, &sc->work[0]); break;
            }
        } else
        switch (axv) {  // element groups per CTA x ring stages (36 KB each): bytes in flight vs. threads per SM
            case 1: launch_ax_cg<8, 2, 3>(r.p, p.p, a.u, ap.p, a.nel, iter == 1, &sc->work[0]); break;
            case 2: launch_ax_cg<8, 2, 2>(r.p, p.p, a.u, ap.p, a.nel, iter == 1, &sc->work[0]); break;
            case 3: launch_ax_cg<8, 3, 2>(r.p, p.p, a.u, ap.p, a.nel, iter == 1, &sc->work[0]); break;   // default before kernel v6
            // warp per element, in-plane contractions on DMMA (ax.cuh kernel v6): warps per SM x 36 KB ring stages per warp
            case 10: launch_ax_cg_mma<3, 2>(r.p, p.p, a.u, ap.p, a.nel, iter == 1, &sc->work[0]); break;
            case 11: launch_ax_cg_mma<2, 3>(r.p, p.p, a.u, ap.p, a.nel, iter == 1, &sc->work[0]); break;
            default: launch_ax_cg_mma<6, 1>(r.p, p.p, a.u, ap.p, a.nel, iter == 1, &sc->work[0]); break;
        }
        prof_end(PROF_AX);
        comm_allreduce_sum(&sc->work[0], 1);
        if (structured) {
            prof_begin(PROF_GS);
            if (h.ngroupsE > 0) {
                gs_gval_kernel<<<blocks_for(h.ngroupsE), 256, 0, s>>>(h.gval.p, ap.p, h.goffE.p, h.gidxE.p, (int)h.ngroupsE);
                NEKB_LAUNCHED();
            }
            if (h.ngroupsS > 0) {
                gs_local_kernel<1><<<blocks_for(h.ngroupsS), 256, 0, s>>>(ap.p, h.goffS.p, h.gidxS.p, (int)h.ngroupsS);
                NEKB_LAUNCHED();
            }
            if (h.nshared > 0 || c.nranks > 1) gs_remote_exchange(h, ap.p, 1);
            prof_end(PROF_GS);
            prof_begin(PROF_UPDATE);
            if (gmode == 5) {
                int g5 = grid;
                if (g5 > c.num_sms * 4) g5 = c.num_sms * 4;
                cggos_update5_kernel<<<g5, CG_THREADS, 0, s>>>(r.p, ap.p, c.wcode.p, h.ftab.p, h.etab.p, h.gval.p, n, sc,
                                                               c.partials.p + 2 * CG_PART_STRIDE);
            } else if (gmode == 6) {
                // warp-per-element TMA ring (cggos_update6_kernel): warps per SM x stages per warp
                static int var6 = -1;
                if (var6 < 0) {
                    const char *e = getenv("NEKB_UPD6_VARIANT");
                    var6 = e ? atoi(e) : 0;
                }
                double *part = c.partials.p + 2 * CG_PART_STRIDE;
                switch (var6) {
                    case 1: launch_cggos_update6<8, 3>(r.p, ap.p, c.wcode.p, h, a.nel, sc, part); break;
                    case 2: launch_cggos_update6<10, 2>(r.p, ap.p, c.wcode.p, h, a.nel, sc, part); break;
                    case 3: launch_cggos_update6<6, 4>(r.p, ap.p, c.wcode.p, h, a.nel, sc, part); break;
                    default: launch_cggos_update6<12, 2>(r.p, ap.p, c.wcode.p, h, a.nel, sc, part); break;
                }
            } else if (gmode == 4) {
                // resident CTAs per SM x elements per CTA = elements in flight per SM (the kernel is latency-bound: ncu
                // profiles/r2g: DRAM 51 % active at 3.49 GB moved, which is within 4 % of the minimum)
                static int var = -1;
                if (var < 0) {
                    const char *e = getenv("NEKB_UPD4_VARIANT");
                    var = e ? atoi(e) : 3;   // measured (profiles/r2i, r2k): one element per 128-thread CTA, 12 CTAs per SM
                }
#define NEKB_UPD4(EPBV, MINBV)                                                                                               \
    cggos_update4_kernel<EPBV, MINBV><<<grid_for((a.nel + EPBV - 1) / EPBV, MINBV), 128 * EPBV, 0, s>>>(                      \
        r.p, ap.p, c.wcode.p, h.ftab.p, h.etab.p, h.gval.p, a.nel, sc, c.partials.p + 2 * CG_PART_STRIDE)
                switch (var) {
                    case 0: NEKB_UPD4(2, 4); break;
                    case 1: NEKB_UPD4(2, 6); break;
                    case 2: NEKB_UPD4(1, 8); break;
                    case 4: NEKB_UPD4(4, 3); break;
                    default: NEKB_UPD4(1, 12); break;
                }
#undef NEKB_UPD4
            } else
                cggos_update3_kernel<<<grid, CG_THREADS, 0, s>>>(r.p, ap.p, c.wcode.p, h.ftab.p, h.etab.p, h.gval.p, n, sc,
                                                                 c.partials.p + 2 * CG_PART_STRIDE);
            NEKB_LAUNCHED();
            prof_end(PROF_UPDATE);
        } else if (gather) {
            if (gmode == 2 && h.ngroups3 > 0) {
                prof_begin(PROF_GS);
                gs_local_kernel<1><<<blocks_for(h.ngroups3), 256, 0, s>>>(ap.p, h.goff3.p, h.gidx3.p, (int)h.ngroups3);
                NEKB_LAUNCHED();
                prof_end(PROF_GS);
            }
            prof_begin(PROF_UPDATE);
            cggos_update2_gs_kernel<<<grid, CG_THREADS, 0, s>>>(r.p, ap.p, c.wcode.p, h.link.p, h.goff.p, h.gidx.p, n, sc,
                                                                c.partials.p + 2 * CG_PART_STRIDE);
            NEKB_LAUNCHED();
            prof_end(PROF_UPDATE);
        } else {
            prof_begin(PROF_GS);
            gs_op(a.gs_handle, ap.p, 1, nullptr);
            prof_end(PROF_GS);
            prof_begin(PROF_UPDATE);
            cggos_update2_kernel<<<grid, CG_THREADS, 0, s>>>(r.p, ap.p, c.wcode.p, n, sc, c.partials.p + 2 * CG_PART_STRIDE);
            NEKB_LAUNCHED();
            prof_end(PROF_UPDATE);
        }
        comm_allreduce_sum(&sc->work[1], 1);
        if (hist_host) {
            cggos_hist_kernel<<<1, 1, 0, s>>>(sc, c.hist.p, iter - 1);
            NEKB_LAUNCHED();
        }
    }
    prof_begin(PROF_PUPDATE);
    axpy_alpha_kernel<<<grid, CG_THREADS, 0, s>>>(a.u, p.p, n, sc);
    NEKB_LAUNCHED();
    prof_end(PROF_PUPDATE);
    if (prof().on) {
        NEKB_CUDA(cudaStreamSynchronize(s));
        prof_collect();
    }
    if (hist_host) {
        std::vector<double> hh((size_t)3 * maxit);
        NEKB_CUDA(cudaMemcpyAsync(hh.data(), c.hist.p, sizeof(double) * hh.size(), cudaMemcpyDeviceToHost, s));
        NEKB_CUDA(cudaStreamSynchronize(s));
        for (size_t q = 0; q < hh.size(); q++) hist_host[q] = hh[q];
    }
    return maxit;
}

// ---------------------------------------------------------------------------------------------- cggo
// core/hmholtz.f:673-679: tol = abs(tin); a non-zero restol(ifield) overrules it; a negative tin (relative tolerance)
// overrules both.  The kernels take "tin" and apply abs() / the relative rule, so the override is folded in here.
inline double cggo_tin(double tin)
{
    Ctx &c = ctx();
    const double rt = (c.ifield >= 0 && c.ifield < 32) ? c.restol[c.ifield] : 0.0;
    return (tin >= 0.0 && rt != 0.0) ? rt : tin;
}
struct CggoArgs {
    double *x;
    const double *f, *h1, *h2, *mask, *mult, *binv;
    int gs_handle;
    int nel;
    double vol;
    int istep;
    bool pres = false;   // name = 'PRES' with param(42) = 1: the plain-PCG pressure solve (hmholtz.f:710-712, :737-751)
};

// hmholtz.f:695-697: r=f, x=0, p=0 ; :699 fmax = glamax(f)
__global__ void __launch_bounds__(CG_THREADS)
    cggo_init_kernel(double *__restrict__ x, double *__restrict__ r, double *__restrict__ p,
                     const double *__restrict__ f, int64_t n, CgScalars *sc, double *partials)
{
    __shared__ double red[33];
    double m = 0.0;
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < n; t += (int64_t)gridDim.x * blockDim.x) {
        const double fv = f[t];
        x[t] = 0.0;
        r[t] = fv;
        p[t] = 0.0;
        m = fmax(m, fabs(fv));
    }
    double b = block_reduce<true>(m, red);
    grid_reduce<true>(b, partials, &sc->counter[1], red, [=](double tot) {
        sc->work[2] = tot;  // fmax
        sc->it = 0;
        sc->done = 0;
        sc->niter = 0;
        sc->rtz1 = 1.0;  // :723
        sc->rho = 0.0;
    });
}

// :730 z = r*d ; :754-760 scalar(1) = sum z r mult, scalar(2) = sum mult binv r r (vlsc32)
__global__ void __launch_bounds__(CG_THREADS)
    cggo_dots_kernel(const double *__restrict__ r, const double *__restrict__ d, const double *__restrict__ mult,
                     const double *__restrict__ binv, int64_t n, CgScalars *sc, double *partials)
{
    __shared__ double red[33];
    if (sc->done) return;
    double s1 = 0.0, s2 = 0.0;
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < n; t += (int64_t)gridDim.x * blockDim.x) {
        const double rv = r[t], mu = mult[t];
        const double z = rv * d[t];
        s1 = fma(z * rv, mu, s1);
        s2 = fma(mu * binv[t] * rv, rv, s2);
    }
    double b1 = block_reduce(s1, red);
    double b2 = block_reduce(s2, red);
    grid_reduce(b1, partials, &sc->counter[2], red, [=](double tot) { sc->work[0] = tot; });
    grid_reduce(b2, partials + CG_PART_STRIDE, &sc->counter[3], red, [=](double tot) { sc->work[1] = tot; });
}

// the same with an explicit preconditioned residual z (Schwarz branch, :737-745)
__global__ void __launch_bounds__(CG_THREADS)
    cggo_dots_z_kernel(const double *__restrict__ r, const double *__restrict__ z, const double *__restrict__ mult,
                       const double *__restrict__ binv, int64_t n, CgScalars *sc, double *partials)
{
    __shared__ double red[33];
    if (sc->done) return;
    double s1 = 0.0, s2 = 0.0;
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < n; t += (int64_t)gridDim.x * blockDim.x) {
        const double rv = r[t], mu = mult[t];
        s1 = fma(z[t] * rv, mu, s1);
        s2 = fma(mu * binv[t] * rv, rv, s2);
    }
    double b1 = block_reduce(s1, red);
    double b2 = block_reduce(s2, red);
    grid_reduce(b1, partials, &sc->counter[2], red, [=](double tot) { sc->work[0] = tot; });
    grid_reduce(b2, partials + CG_PART_STRIDE, &sc->counter[3], red, [=](double tot) { sc->work[1] = tot; });
}
__global__ void __launch_bounds__(CG_THREADS)
    cggo_pupdate_z_kernel(double *__restrict__ p, const double *__restrict__ z, int64_t n, const CgScalars *sc)
{
    if (sc->done) return;
    const double beta = (sc->it == 0) ? 0.0 : sc->rtz1 / sc->rtz2;
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < n; t += (int64_t)gridDim.x * blockDim.x)
        p[t] = fma(beta, p[t], z[t]);
}

// z = r * d (explicit, for the null-space correction) ; a += s[0]*s[1]*b ; a += s[0]*s[1]
__global__ void __launch_bounds__(CG_THREADS)
    cggo_z_kernel(double *__restrict__ z, const double *__restrict__ r, const double *__restrict__ d, int64_t n, const CgScalars *sc)
{
    if (sc->done) return;
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < n; t += (int64_t)gridDim.x * blockDim.x) z[t] = r[t] * d[t];
}
__global__ void __launch_bounds__(CG_THREADS)
    cggo_sum2_kernel(const double *__restrict__ a, const double *__restrict__ b, int64_t n, double *out, const CgScalars *sc,
                     double *partials, unsigned *counter)
{
    __shared__ double red[33];
    if (sc->done) return;
    double s = 0.0;
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < n; t += (int64_t)gridDim.x * blockDim.x) s = fma(a[t], b ? b[t] : 1.0, s);
    const double bs = block_reduce(s, red);
    grid_reduce(bs, partials, counter, red, [=](double tot) { *out = tot; });
}
// a += smean * (*dot) * (b ? b : 1)   (hmholtz.f:713-718 with b = dssum(bm1), :750-751 with b absent)
__global__ void __launch_bounds__(CG_THREADS)
    cggo_mcor_kernel(double *__restrict__ a, const double *__restrict__ b, double smean, const double *dot, int64_t n, const CgScalars *sc)
{
    if (sc->done) return;
    const double rmean = smean * *dot;
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < n; t += (int64_t)gridDim.x * blockDim.x)
        a[t] = b ? fma(rmean, b[t], a[t]) : a[t] + rmean;
}
__global__ void __launch_bounds__(CG_THREADS)
    negmax_kernel(const double *__restrict__ a, int64_t n, double *out, double *partials, unsigned *counter)
{
    __shared__ double red[33];
    double m = -1.0e300;
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < n; t += (int64_t)gridDim.x * blockDim.x) m = fmax(m, -a[t]);
    const double b = block_reduce<true>(m, red);
    grid_reduce<true>(b, partials, counter, red, [=](double tot) { *out = tot; });  // max(-a) = -min a
}

// :761-791 scalar bookkeeping and the convergence test (single thread).
__global__ void cggo_check_kernel(CgScalars *sc, double vol, double tin, int istep, int niter_max, double *hist, double param22)
{
    if (sc->done) return;
    const int iter = sc->it + 1;
    sc->rtz2 = sc->rtz1;
    sc->rtz1 = sc->work[0];
    const double rbn2 = sqrt(sc->work[1] / vol);
    sc->rbn2 = rbn2;
    if (iter == 1) {
        sc->rbn0 = rbn2;
        double tol = fabs(tin);                               // :673
        if (param22 < 0) tol = fabs(param22) * rbn2;          // :764
        if (tin < 0) tol = fabs(tin) * rbn2;                  // :765
        sc->tol = tol;
    }
    hist[3 * (iter - 1) + 0] = sc->rtz1;
    hist[3 * (iter - 1) + 1] = rbn2;
    if (rbn2 <= sc->tol && (iter > 1 || istep <= 5)) {  // :778
        sc->done = 1;
        sc->niter = iter - 1;
    } else if (iter > niter_max) {
        sc->done = 1;
        sc->niter = niter_max;
    }
}

// :793-795 beta = rtz1/rtz2 (0 on the first iteration) ; p = z + beta p  (add2s1)
__global__ void __launch_bounds__(CG_THREADS)
    cggo_pupdate_kernel(double *__restrict__ p, const double *__restrict__ r, const double *__restrict__ d, int64_t n,
                        const CgScalars *sc)
{
    if (sc->done) return;
    const double beta = (sc->it == 0) ? 0.0 : sc->rtz1 / sc->rtz2;
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < n; t += (int64_t)gridDim.x * blockDim.x)
        p[t] = fma(beta, p[t], r[t] * d[t]);
}

// :798 w *= mask ; :801 rho = glsc3(w,p,mult)
__global__ void __launch_bounds__(CG_THREADS)
    cggo_rho_kernel(double *__restrict__ w, const double *__restrict__ p, const double *__restrict__ mask,
                    const double *__restrict__ mult, int64_t n, CgScalars *sc, double *partials)
{
    __shared__ double red[33];
    if (sc->done) return;
    double s = 0.0;
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < n; t += (int64_t)gridDim.x * blockDim.x) {
        const double wv = w[t] * mask[t];
        w[t] = wv;
        s = fma(wv * p[t], mult[t], s);
    }
    double b = block_reduce(s, red);
    grid_reduce(b, partials, &sc->counter[2], red, [=](double tot) { sc->rho = tot; });
}

// :802-805 alpha = rtz1/rho ; x += alpha p ; r -= alpha w.  Last block advances the counter.
__global__ void __launch_bounds__(CG_THREADS)
    cggo_xr_kernel(double *__restrict__ x, double *__restrict__ r, const double *__restrict__ p,
                   const double *__restrict__ w, int64_t n, CgScalars *sc, double *hist)
{
    if (sc->done) return;
    const double rho = sc->rho;
    const double alpha = sc->rtz1 / rho;
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < n; t += (int64_t)gridDim.x * blockDim.x) {
        x[t] = fma(alpha, p[t], x[t]);
        r[t] = fma(-alpha, w[t], r[t]);
    }
    __shared__ int s_last;
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        unsigned t = atomicInc(&sc->counter[1], gridDim.x - 1);
        s_last = (t == gridDim.x - 1);
    }
    __syncthreads();
    if (s_last && threadIdx.x == 0) {
        hist[3 * sc->it + 2] = rho;
        sc->it = sc->it + 1;
    }
}

__global__ void __launch_bounds__(CG_THREADS)
    invcol1_kernel(double *__restrict__ a, int64_t n)
{
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < n; t += (int64_t)gridDim.x * blockDim.x)
        a[t] = 1.0 / a[t];
}

__global__ void __launch_bounds__(CG_THREADS)
    absmax_kernel(const double *__restrict__ a, int64_t n, double *out, double *partials, unsigned *counter)
{
    __shared__ double red[33];
    double m = 0.0;
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < n; t += (int64_t)gridDim.x * blockDim.x)
        m = fmax(m, fabs(a[t]));
    double b = block_reduce<true>(m, red);
    grid_reduce<true>(b, partials, counter, red, [=](double tot) { *out = tot; });
}

// dpc = 1 / dssum(setprec_local)   (hmholtz.f:380-524)
inline void setprec_run(double *dpc, const double *h1, const double *h2, int nel, int gs_handle)
{
    Ctx &c = ctx();
    const int64_t n = (int64_t)nel * c.nxyz;
    launch_setprec(dpc, h1, h2, nel);
    gs_op(gs_handle, dpc, 1, nullptr);
    invcol1_kernel<<<cg_grid(n), CG_THREADS, 0, c.stream>>>(dpc, n);
    NEKB_LAUNCHED();
}

// Returns niterhm.  hist_host (may be NULL): 3 doubles per executed iteration (rtz1, rbn2, rho).
// Includes the all-Neumann null-space correction (ifmcor, :705-720, :749-752).  The 'PRES' branch (:641-657) is taken by
// the callers (cggo_, hmholtz_, hsolve_ forward to hmh_gmres / hmh_flex_cg for param(42) = 0 / 2); with param(42) = 1 they come
// here with a.pres set: no ifmcor (:710-711), Schwarz + coarse-grid correction z = fdm_h1(r) + crs_solve_h1(r) (:741-744) or
// Jacobi (kfldfdm < 0), and ortho(z) (:747-748) in every iteration.
// lx1 = 8 Jacobi solves normally go through the fused path (hcg.cuh cggo_solve); this routine is the general one.
inline int cggo_run(const CggoArgs &a, double tin, int maxit, double *hist_host)
{
    Ctx &c = ctx();
    const int64_t n = (int64_t)a.nel * c.nxyz;
    cudaStream_t s = c.stream;
    const int maxcg = 900;
    const int niter = maxit < maxcg ? maxit : maxcg;
    DevBuf<double> &r = c.work[0], &p = c.work[1], &w = c.work[2], &d = c.work[3];
    r.ensure(n);
    p.ensure(n);
    w.ensure(n);
    d.ensure(n);
    c.hist.ensure((size_t)3 * (niter + 2));
    NEKB_CUDA(cudaMemsetAsync(c.hist.p, 0, sizeof(double) * 3 * (niter + 2), s));
    c.partials.ensure(4 * CG_PART_STRIDE);
    CgScalars *sc = c.sc.p;
    const int grid = cg_grid(n);
    GsMap &h = gs_get(a.gs_handle);
    NEKB_REQUIRE(h.n == n, "cggo: gs handle was set up for a different vector length");

    // ifh2 (setfast :303-305): max|h2| > 0
    absmax_kernel<<<grid, CG_THREADS, 0, s>>>(a.h2, n, &sc->work[3], c.partials.p, &sc->counter[0]);
    NEKB_LAUNCHED();
    comm_allreduce_max(&sc->work[3], 1);
    double h2max = 0.0;
    NEKB_CUDA(cudaMemcpyAsync(&h2max, &sc->work[3], sizeof(double), cudaMemcpyDeviceToHost, s));
    NEKB_CUDA(cudaStreamSynchronize(s));
    const double *h2_eff = (h2max > 0.0) ? a.h2 : nullptr;

    const bool schwarz = fdm_h1_kfldfdm() >= 0;  // :686-691
    DevBuf<double> &z = c.work[4];
    if (schwarz) {
        z.ensure(n);
        set_fdm_prec_h1b_dev(d.p, a.h1, a.h2, a.nel);
    } else
        setprec_run(d.p, a.h1, a.h2, a.nel, a.gs_handle);  // :690
    cggo_init_kernel<<<grid, CG_THREADS, 0, s>>>(a.x, r.p, p.p, a.f, n, sc, c.partials.p + 1 * CG_PART_STRIDE);
    NEKB_LAUNCHED();
    comm_allreduce_max(&sc->work[2], 1);
    double fmax = 0.0;
    NEKB_CUDA(cudaMemcpyAsync(&fmax, &sc->work[2], sizeof(double), cudaMemcpyDeviceToHost, s));
    NEKB_CUDA(cudaStreamSynchronize(s));
    if (fmax == 0.0) return 0;  // :700-701

    // :705-720 non-trivial null space: h2 = 0 everywhere and no Dirichlet node
    negmax_kernel<<<grid, CG_THREADS, 0, s>>>(a.mask, n, &sc->work[3], c.partials.p, &sc->counter[0]);
    NEKB_LAUNCHED();
    comm_allreduce_max(&sc->work[3], 1);
    double skmin = 0.0;
    NEKB_CUDA(cudaMemcpyAsync(&skmin, &sc->work[3], sizeof(double), cudaMemcpyDeviceToHost, s));
    NEKB_CUDA(cudaStreamSynchronize(s));
    skmin = -skmin;  // glmin(mask)
    const bool ifmcor = !a.pres && skmin > 0.0 && h2max == 0.0;   // :710-712: 'PRES' skips the mean correction of r
    double smean = 0.0;
    DevBuf<double> &bsum = c.work[5];
    const bool explicit_z = schwarz || ifmcor || a.pres;
    if (a.pres) z.ensure(n);
    if (ifmcor) {
        NEKB_REQUIRE(c.bm1.n >= (size_t)n, "cggo: bm1 must be registered for the null-space correction");
        z.ensure(n), bsum.ensure(n);
        cggo_sum2_kernel<<<grid, CG_THREADS, 0, s>>>(c.bm1.p, nullptr, n, &sc->work[3], sc, c.partials.p, &sc->counter[0]);
        NEKB_LAUNCHED();
        comm_allreduce_sum(&sc->work[3], 1);
        double vsum = 0.0;
        NEKB_CUDA(cudaMemcpyAsync(&vsum, &sc->work[3], sizeof(double), cudaMemcpyDeviceToHost, s));
        NEKB_CUDA(cudaStreamSynchronize(s));
        smean = -1.0 / vsum;                                                   // :711
        cggo_sum2_kernel<<<grid, CG_THREADS, 0, s>>>(r.p, a.mult, n, &sc->work[3], sc, c.partials.p, &sc->counter[0]);  // :712
        NEKB_LAUNCHED();
        comm_allreduce_sum(&sc->work[3], 1);
        NEKB_CUDA(cudaMemcpyAsync(bsum.p, c.bm1.p, sizeof(double) * (size_t)n, cudaMemcpyDeviceToDevice, s));
        gs_op(a.gs_handle, bsum.p, 1, nullptr);                                // :713-714
        cggo_mcor_kernel<<<grid, CG_THREADS, 0, s>>>(r.p, bsum.p, smean, &sc->work[3], n, sc);  // :715
        NEKB_LAUNCHED();
    }

    int launched = 0, result = -1;
    const int batch = 8;  // iterations enqueued between two looks at the convergence flag
    while (result < 0) {
        for (int b = 0; b < batch; b++) {
            if (explicit_z) {
                if (schwarz) {  // :737-745
                    fdm_h1_apply(z.p, r.p, d.p, a.mask, a.nel, a.gs_handle);
                    if (a.pres) {   // :742-744 crs_solve_h1(w,r) ; z += w  ("currently, crs grd only for P")
                        cggo_pres_coarse(w.p, r.p, a.mult, a.nel);
                        add2_kernel<<<grid, CG_THREADS, 0, s>>>(z.p, w.p, n);
                        NEKB_LAUNCHED();
                    }
                } else {
                    cggo_z_kernel<<<grid, CG_THREADS, 0, s>>>(z.p, r.p, d.p, n, sc);
                    NEKB_LAUNCHED();
                }
                if (a.pres) cggo_pres_ortho(z.p, n);   // :747-748
                if (ifmcor) {  // :749-752 rmean = smean*glsc2(z,bm1) ; z += rmean
                    cggo_sum2_kernel<<<grid, CG_THREADS, 0, s>>>(z.p, c.bm1.p, n, &sc->work[3], sc, c.partials.p, &sc->counter[0]);
                    NEKB_LAUNCHED();
                    comm_allreduce_sum(&sc->work[3], 1);
                    cggo_mcor_kernel<<<grid, CG_THREADS, 0, s>>>(z.p, nullptr, smean, &sc->work[3], n, sc);
                    NEKB_LAUNCHED();
                }
                cggo_dots_z_kernel<<<grid, CG_THREADS, 0, s>>>(r.p, z.p, a.mult, a.binv, n, sc, c.partials.p + 2 * CG_PART_STRIDE);
            } else
                cggo_dots_kernel<<<grid, CG_THREADS, 0, s>>>(r.p, d.p, a.mult, a.binv, n, sc, c.partials.p + 2 * CG_PART_STRIDE);
            NEKB_LAUNCHED();
            comm_allreduce_sum(&sc->work[0], 2);
            cggo_check_kernel<<<1, 1, 0, s>>>(sc, a.vol, tin, a.istep, niter, c.hist.p, c.param[22]);
            NEKB_LAUNCHED();
            if (explicit_z)
                cggo_pupdate_z_kernel<<<grid, CG_THREADS, 0, s>>>(p.p, z.p, n, sc);
            else
                cggo_pupdate_kernel<<<grid, CG_THREADS, 0, s>>>(p.p, r.p, d.p, n, sc);
            NEKB_LAUNCHED();
            launch_ax(p.p, w.p, a.h1, h2_eff, a.nel, nullptr);
            gs_op(a.gs_handle, w.p, 1, nullptr);
            cggo_rho_kernel<<<grid, CG_THREADS, 0, s>>>(w.p, p.p, a.mask, a.mult, n, sc, c.partials.p + 2 * CG_PART_STRIDE);
            NEKB_LAUNCHED();
            comm_allreduce_sum(&sc->rho, 1);
            cggo_xr_kernel<<<grid, CG_THREADS, 0, s>>>(a.x, r.p, p.p, w.p, n, sc, c.hist.p);
            NEKB_LAUNCHED();
            launched++;
        }
        CgScalars hs;
        NEKB_CUDA(cudaMemcpyAsync(&hs, sc, sizeof(CgScalars), cudaMemcpyDeviceToHost, s));
        NEKB_CUDA(cudaStreamSynchronize(s));
        if (hs.done) result = hs.niter;
        NEKB_REQUIRE(launched <= niter + 2 * batch, "cggo: convergence flag never raised");
    }
    if (hist_host) {
        std::vector<double> hh((size_t)3 * (niter + 2));
        NEKB_CUDA(cudaMemcpyAsync(hh.data(), c.hist.p, sizeof(double) * hh.size(), cudaMemcpyDeviceToHost, s));
        NEKB_CUDA(cudaStreamSynchronize(s));
        const int rows = result + 1 <= niter + 1 ? result + 1 : niter + 1;
        for (int i = 0; i < 3 * rows; i++) hist_host[i] = hh[i];
    }
    return result;
}

}  // namespace nekb
