// hcg.cuh -- fused, multi-right-hand-side Jacobi-PCG for the Helmholtz solves.
//
// What it replaces: the three consecutive `hsolve -> hmholtz -> cggo` calls of ophinv (core/induct.f:1022-1090; VELX, VELY,
// VELZ share h1, h2, the geometric factors, the Jacobi diagonal, mult and binvm1 and differ only in right-hand side, mask
// and scalars), and -- with one right-hand side -- the Jacobi branch of cggo itself (core/hmholtz.f:611-846).  Every
// component runs exactly the reference's recurrence with its own scalars and its own exit test (:778), so iteration counts
// and iterates are those of three independent solves; a converged component simply stops being touched.
//
// Why: the stock sequence streams 30 words per grid point per iteration and component (dots 4, p-update 4, Ax 9, gs 2.4,
// rho 5, x/r-update 6).  Here one pass over an element's tiles does, for all components at once,
//     x_c += alpha_c p_c ;  z = d r_c ;  p_c = z + beta_c p_c ;  w_c = h1 D^T G D p_c + (h2 B) p_c
// with the six factor tiles, h1, h2*B and d fetched ONCE per element (cp.async.bulk + mbarrier ring, as ax_cg_kernel), masks
// are packed to one byte per node for all components, and the residual update is fused with the two dot products of the
// next iteration:    front 6 + 9/NRHS, gs 2.4, update 3 + 2.1/NRHS words  ->  15.1 words per component at NRHS = 3 (22.5 at
// NRHS = 1) instead of 30; rho = (w,p) is summed inside the front kernel from the un-assembled w (round 1 spent a pass of
// 2 + 1.1/NRHS words on it).
// Tried and rejected on measurement (round 2, profiles/r2y_ophinv*.json): the structured gather of cggos_update6_kernel applied
// here (one warp per element behind a TMA ring of r_c, w_c, mult, d, binv stages, face partners / edge values gathered in the
// update): a 3-component stage is 37 KB, so only 3 warps fit an SM and the kernel ran 2x slower than gs_op + hcg_update_kernel
// (6.9 vs 3.2 ms per iteration and component); with one component (5 warps) it was a draw (4.03 vs 4.00 ms).
#pragma once
#include "cg.cuh"

namespace nekb {

constexpr int HCG_MAXR = 3;

struct HcgComp {
    double rtz1, rtz2, rho, alpha, beta, rbn2, rbn0, tol;
    double work[2];        // (z,r)_mult and (r,r)_mult,binv deposited by the update kernel
    int it, done, niter, pending;  // pending: x += alpha p of the last completed iteration not applied yet
};
struct HcgScalars {
    HcgComp c[HCG_MAXR];
    unsigned counter[8];
    int alldone, pad;
};
struct HcgPtrs {
    double *x[HCG_MAXR], *r[HCG_MAXR], *p[HCG_MAXR], *w[HCG_MAXR];
};

template <int NX, int NRHS, int PIPES, int STAGES>
struct HcgSmem {
    static constexpr int N2 = NX * NX, N3 = NX * NX * NX;
    static constexpr int shared_tiles = 9;                                         // 6 factors, h1, h2*B, d
    static constexpr size_t stage_doubles = (size_t)(shared_tiles + 3 * NRHS) * N3;  // + r, p, x per component
    static constexpr size_t pipe_doubles = STAGES * stage_doubles;
    static constexpr size_t bytes = PIPES * pipe_doubles * sizeof(double) + PIPES * STAGES * sizeof(uint64_t) + 64 * sizeof(double);
};

// Front half of the iteration.  CTA = PIPES element pipelines x NRHS component groups x NX*NX threads; thread (i,j) of a
// group owns the k-column of its component.  The groups of a pipeline share the stage (one TMA fill per element).
template <int NX, int NRHS, int PIPES, int STAGES, bool HAS_H2>
__global__ void __launch_bounds__(NX *NX *NRHS *PIPES, 1)
    hcg_front_kernel(HcgPtrs P, const double *__restrict__ g, const double *__restrict__ h1, const double *__restrict__ h2b,
                     const double *__restrict__ d, int nel, HcgScalars *sc, double *partials)
{
    using L = HcgSmem<NX, NRHS, PIPES, STAGES>;
    constexpr int N2 = L::N2, N3 = L::N3, GT = N2 * NRHS;
    constexpr uint32_t T_BYTES = N3 * sizeof(double);
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double *smem = reinterpret_cast<double *>(smem_raw);
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + PIPES * L::pipe_doubles);
    double *s_red = reinterpret_cast<double *>(bars + PIPES * STAGES);   // 64 doubles: [0,16) warp sums, [16,64) reduction scratch

    const int pipe = threadIdx.x / GT, tp = threadIdx.x % GT;
    const int comp = tp / N2, ij = tp % N2, i = ij % NX, j = ij / NX;
    double *pbase = smem + pipe * L::pipe_doubles;
    uint64_t *full = bars + pipe * STAGES;
    const bool leader = tp == 0;

    if (threadIdx.x == 0) {
        for (int q = 0; q < PIPES * STAGES; q++) mbar_init(&bars[q], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    bool act[NRHS];
    int nact = 0;
#pragma unroll
    for (int c = 0; c < NRHS; c++) {
        act[c] = sc->c[c].done == 0;
        nact += act[c] ? 1 : 0;
    }
    if (nact == 0) return;  // every component has converged (launches enqueued past the exit test)
    const bool mine = act[comp];
    const double alpha = (mine && sc->c[comp].pending) ? sc->c[comp].alpha : 0.0;
    const double beta = mine ? sc->c[comp].beta : 0.0;
    const bool upd_x = mine && sc->c[comp].pending;

    double Di[NX], Dj[NX], DTi[NX], DTj[NX];
#pragma unroll
    for (int m = 0; m < NX; m++) {
        Di[m] = c_D[i * NX + m];
        Dj[m] = c_D[j * NX + m];
        DTi[m] = c_D[m * NX + i];
        DTj[m] = c_D[m * NX + j];
    }

    const int first = blockIdx.x * PIPES + pipe, stride = gridDim.x * PIPES;
    auto issue = [&](int stage, int e) {
        double *st = pbase + stage * L::stage_doubles;
        mbar_expect_tx(&full[stage], (uint32_t)((8 + (HAS_H2 ? 1 : 0) + 3 * nact) * T_BYTES));
        bulk_g2s(st, g + (size_t)e * 6 * N3, 6 * T_BYTES, &full[stage]);
        bulk_g2s(st + 6 * N3, h1 + (size_t)e * N3, T_BYTES, &full[stage]);
        if (HAS_H2) bulk_g2s(st + 7 * N3, h2b + (size_t)e * N3, T_BYTES, &full[stage]);
        bulk_g2s(st + 8 * N3, d + (size_t)e * N3, T_BYTES, &full[stage]);
#pragma unroll
        for (int c = 0; c < NRHS; c++)
            if (act[c]) {
                double *sc_ = st + (9 + 3 * c) * N3;
                bulk_g2s(sc_, P.r[c] + (size_t)e * N3, T_BYTES, &full[stage]);
                bulk_g2s(sc_ + N3, P.p[c] + (size_t)e * N3, T_BYTES, &full[stage]);
                bulk_g2s(sc_ + 2 * N3, P.x[c] + (size_t)e * N3, T_BYTES, &full[stage]);
            }
    };
    if (leader) {
#pragma unroll
        for (int s = 0; s < STAGES; s++) {
            const int e = first + s * stride;
            if (e < nel) issue(s, e);
        }
    }

    double *__restrict__ pg = P.p[comp], *__restrict__ xg = P.x[comp], *__restrict__ wg = P.w[comp];
    // rho_c = (w_c, p_c)_{mask mult} of hmholtz.f:798-801 is accumulated here from the UN-assembled w: p is continuous and zero on
    // masked nodes (p = d r + beta p with r, d assembled and masked), so sum_nodes mult*mask*dssum(w)*p = sum_nodes w_local*p --
    // the identity bp5.usr's own ax_e_bp5 uses for pap.  This replaces a pass over w, p, mult and the mask bytes per iteration.
    double rho_loc = 0.0;
    int it = 0;
    for (int e = first; e < nel; e += stride, it++) {
        const int stage = it % STAGES;
        const uint32_t parity = (uint32_t)(it / STAGES) & 1u;
        mbar_wait(&full[stage], parity);
        double *sg = pbase + stage * L::stage_doubles;
        const double *sh1 = sg + 6 * N3, *sh2 = sg + 7 * N3, *sd = sg + 8 * N3;
        double *sr = sg + (9 + 3 * comp) * N3, *sp = sr + N3, *sx = sr + 2 * N3;
        const size_t eo = (size_t)e * N3;
        if (mine) {
            double ucol[NX], wcol[NX];
#pragma unroll
            for (int k = 0; k < NX; k++) {
                const int q = k * N2 + ij;
                const double po = sp[q];
                const double pn = fma(beta, po, sd[q] * sr[q]);   // hmholtz.f:730 z = r*d ; :795 p = z + beta p
                if (upd_x) xg[eo + q] = fma(alpha, po, sx[q]);    // :804 of the previous iteration
                sp[q] = pn;
                pg[eo + q] = pn;
                ucol[k] = pn;
                wcol[k] = 0.0;
            }
            group_barrier(1 + pipe * NRHS + comp, N2);  // this component's p tile is complete
            // the r and x tiles are dead now: they take the r- and s-fluxes of all planes (one barrier per element)
            double *swr = sr, *sws = sx;
#pragma unroll
            for (int k = 0; k < NX; k++) {
                const int q = k * N2 + ij;
                const double G0 = sg[0 * N3 + q], G1 = sg[1 * N3 + q], G2 = sg[2 * N3 + q], G3 = sg[3 * N3 + q],
                             G4 = sg[4 * N3 + q], G5 = sg[5 * N3 + q], hh = sh1[q];
                double ur = 0.0, us = 0.0, ut = 0.0;
#pragma unroll
                for (int m = 0; m < NX; m++) {
                    ur = fma(Di[m], sp[k * N2 + j * NX + m], ur);
                    us = fma(Dj[m], sp[k * N2 + m * NX + i], us);
                    ut = fma(c_D[k * NX + m], ucol[m], ut);
                }
                const double wr = fma(G0, ur, fma(G1, us, G2 * ut)) * hh;   // :200-211 + col2(.,helm1)
                const double ws = fma(G1, ur, fma(G3, us, G4 * ut)) * hh;
                const double wt = fma(G2, ur, fma(G4, us, G5 * ut)) * hh;
                swr[q] = wr;
                sws[q] = ws;
#pragma unroll
                for (int m = 0; m < NX; m++) wcol[m] = fma(c_D[k * NX + m], wt, wcol[m]);
            }
            group_barrier(1 + pipe * NRHS + comp, N2);
#pragma unroll
            for (int k = 0; k < NX; k++) {
                double acc = wcol[k];
#pragma unroll
                for (int m = 0; m < NX; m++) {
                    acc = fma(DTi[m], swr[k * N2 + j * NX + m], acc);
                    acc = fma(DTj[m], sws[k * N2 + m * NX + i], acc);
                }
                if (HAS_H2) acc = fma(sh2[k * N2 + ij], ucol[k], acc);   // :225 addcol4(au,helm2,bm1,u)
                wg[eo + k * N2 + ij] = acc;
                rho_loc = fma(ucol[k], acc, rho_loc);
            }
        }
        // every group of the pipeline is past its last read of the stage before the TMA engine may refill it
        group_barrier(1 + PIPES * NRHS + pipe, GT);
        if (leader) {
            const int en = e + STAGES * stride;
            if (en < nel) {
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                issue(stage, en);
            }
        }
    }
    // per-component sums in a fixed order: warp -> CTA (warps of the component in index order) -> grid (CTAs in index order)
    {
        const double ws = warp_sum(rho_loc);
        __syncthreads();
        if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = ws;
        __syncthreads();
        constexpr int WPG = N2 / 32;   // warps per component group
#pragma unroll
        for (int c = 0; c < NRHS; c++) {
            if (!act[c]) continue;     // uniform over the grid
            double b = 0.0;
            if (threadIdx.x == 0)
                for (int pp = 0; pp < PIPES; pp++)
                    for (int q = 0; q < WPG; q++) b += s_red[(pp * NRHS + c) * WPG + q];
            HcgComp *hc = &sc->c[c];
            grid_reduce(b, partials + c * CG_PART_STRIDE, &sc->counter[c], s_red + 16, [=](double tot) { hc->rho = tot; });
        }
    }
}

// rho_c = (w_c, p_c)_{mask_c * mult}   (hmholtz.f:798-801; the mask is applied where w is consumed, w is not rewritten)
template <int NRHS>
__global__ void __launch_bounds__(CG_THREADS, 4)   // 4 CTAs per SM resident: the grid is sized for exactly that (hcg_run)
    hcg_rho_kernel(HcgPtrs P, const unsigned char *__restrict__ mcode, const double *__restrict__ mult, int64_t n, HcgScalars *sc,
                   double *partials)
{
    __shared__ double red[33];
    bool act[NRHS];
    double s[NRHS];
#pragma unroll
    for (int c = 0; c < NRHS; c++) act[c] = sc->c[c].done == 0, s[c] = 0.0;
    const int64_t n2 = n >> 1;
    const double2 *__restrict__ mult2 = reinterpret_cast<const double2 *>(mult);
    const uchar2 *__restrict__ mc2 = reinterpret_cast<const uchar2 *>(mcode);
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < n2; t += (int64_t)gridDim.x * blockDim.x) {
        const uchar2 mc = mc2[t];
        const double2 m = mult2[t];
#pragma unroll
        for (int c = 0; c < NRHS; c++)
            if (act[c]) {
                const double2 w = reinterpret_cast<const double2 *>(P.w[c])[t], p = reinterpret_cast<const double2 *>(P.p[c])[t];
                const double mx = ((mc.x >> c) & 1u) ? m.x : 0.0, my = ((mc.y >> c) & 1u) ? m.y : 0.0;
                s[c] = fma(w.x * p.x, mx, s[c]);
                s[c] = fma(w.y * p.y, my, s[c]);
            }
    }
#pragma unroll
    for (int c = 0; c < NRHS; c++) {
        if (!act[c]) continue;
        const double b = block_reduce(s[c], red);
        HcgComp *hc = &sc->c[c];
        grid_reduce(b, partials + c * CG_PART_STRIDE, &sc->counter[c], red, [=](double tot) { hc->rho = tot; });
    }
}

// alpha = rtz1/rho ; r -= alpha mask w (:802-805) fused with the two sums that open the next iteration (:755-760):
// (z,r)_mult with z = d r, and (r,r)_{mult binv}.  FIRST: r is the right-hand side, nothing to subtract.
// Two nodes per thread and step (16-byte loads, all loads of a step independent of the mask byte).
template <int NRHS, bool FIRST>
__global__ void __launch_bounds__(CG_THREADS, 4)   // 4 CTAs per SM resident: the grid is sized for exactly that (hcg_run)
    hcg_update_kernel(HcgPtrs P, const unsigned char *__restrict__ mcode, const double *__restrict__ mult, const double *__restrict__ d,
                      const double *__restrict__ binv, int64_t n, HcgScalars *sc, double *partials, double *hist, int hist_stride)
{
    __shared__ double red[33];
    bool act[NRHS];
    double al[NRHS], s1[NRHS], s2[NRHS];
#pragma unroll
    for (int c = 0; c < NRHS; c++) {
        act[c] = sc->c[c].done == 0;
        al[c] = (!FIRST && act[c]) ? sc->c[c].rtz1 / sc->c[c].rho : 0.0;
        s1[c] = s2[c] = 0.0;
    }
    const int64_t n2 = n >> 1;  // n = 512 * nel is even
    const double2 *__restrict__ mult2 = reinterpret_cast<const double2 *>(mult), *__restrict__ d2 = reinterpret_cast<const double2 *>(d),
                               *__restrict__ binv2 = reinterpret_cast<const double2 *>(binv);
    const uchar2 *__restrict__ mc2 = reinterpret_cast<const uchar2 *>(mcode);
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < n2; t += (int64_t)gridDim.x * blockDim.x) {
        const uchar2 mc = mc2[t];
        const double2 m = mult2[t], dd = d2[t], bi = binv2[t];
        double2 rv[NRHS], wv[NRHS];
#pragma unroll
        for (int c = 0; c < NRHS; c++)
            if (act[c]) {
                rv[c] = reinterpret_cast<const double2 *>(P.r[c])[t];
                if (!FIRST) wv[c] = reinterpret_cast<const double2 *>(P.w[c])[t];
            }
#pragma unroll
        for (int c = 0; c < NRHS; c++) {
            if (!act[c]) continue;
            double2 r = rv[c];
            if (!FIRST) {
                const double rx = fma(-al[c], wv[c].x, r.x), ry = fma(-al[c], wv[c].y, r.y);
                r.x = ((mc.x >> c) & 1u) ? rx : r.x;
                r.y = ((mc.y >> c) & 1u) ? ry : r.y;
                reinterpret_cast<double2 *>(P.r[c])[t] = r;
            }
            s1[c] = fma(r.x * dd.x * r.x, m.x, s1[c]);
            s1[c] = fma(r.y * dd.y * r.y, m.y, s1[c]);
            s2[c] = fma(r.x * r.x * m.x, bi.x, s2[c]);
            s2[c] = fma(r.y * r.y * m.y, bi.y, s2[c]);
        }
    }
#pragma unroll
    for (int c = 0; c < NRHS; c++) {
        if (!act[c]) continue;
        HcgComp *hc = &sc->c[c];
        const double a = al[c];
        double b = block_reduce(s1[c], red);
        grid_reduce(b, partials + (2 * c) * CG_PART_STRIDE, &sc->counter[c], red, [=](double tot) {
            hc->work[0] = tot;
            if (!FIRST) {
                hc->alpha = a, hc->pending = 1;
                if (hist && 3 * hc->it <= hist_stride) hist[c * hist_stride + 3 * (hc->it - 1) + 2] = hc->rho;
            }
        });
        b = block_reduce(s2[c], red);
        grid_reduce(b, partials + (2 * c + 1) * CG_PART_STRIDE, &sc->counter[3 + c], red, [=](double tot) { hc->work[1] = tot; });
    }
}

// hmholtz.f:755-795 scalar bookkeeping of every component: norms, tolerance overrides, the exit test (:778), beta.
__global__ void hcg_check_kernel(HcgScalars *sc, int nrhs, double vol, double tin0, double tin1, double tin2, int istep, int niter_max,
                                 double param22, double *hist, int hist_stride)
{
    const double tins[3] = {tin0, tin1, tin2};
    int all = 1;
    for (int c = 0; c < nrhs; c++) {
        HcgComp *h = &sc->c[c];
        if (!h->done) {
            const int iter = h->it + 1;
            h->rtz2 = h->rtz1;
            h->rtz1 = h->work[0];
            const double rbn2 = sqrt(h->work[1] / vol);
            h->rbn2 = rbn2;
            if (iter == 1) {
                h->rbn0 = rbn2;
                double tol = fabs(tins[c]);
                if (param22 < 0) tol = fabs(param22) * rbn2;
                if (tins[c] < 0) tol = fabs(tins[c]) * rbn2;
                h->tol = tol;
            }
            if (hist && 3 * iter <= hist_stride) {
                hist[c * hist_stride + 3 * (iter - 1) + 0] = h->rtz1;
                hist[c * hist_stride + 3 * (iter - 1) + 1] = rbn2;
            }
            if (rbn2 <= h->tol && (iter > 1 || istep <= 5)) {
                h->done = 1;
                h->niter = iter - 1;
            } else if (iter > niter_max) {
                h->done = 1;
                h->niter = niter_max;
            } else {
                h->beta = iter == 1 ? 0.0 : h->rtz1 / h->rtz2;
                h->it = iter;
            }
        }
        all &= h->done;
    }
    sc->alldone = all;
}

// x += alpha p of the last completed iteration of every component (the front kernel applies it one iteration late)
template <int NRHS>
__global__ void __launch_bounds__(CG_THREADS) hcg_flush_kernel(HcgPtrs P, int64_t n, HcgScalars *sc)
{
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < n; t += (int64_t)gridDim.x * blockDim.x) {
#pragma unroll
        for (int c = 0; c < NRHS; c++)
            if (sc->c[c].pending) P.x[c][t] = fma(sc->c[c].alpha, P.p[c][t], P.x[c][t]);
    }
}

__global__ void __launch_bounds__(CG_THREADS)
    hcg_prep_kernel(unsigned char *__restrict__ mcode, double *__restrict__ h2b, const double *m0, const double *m1, const double *m2,
                    const double *__restrict__ h2, const double *__restrict__ bm1, int64_t n, int *notbinary)
{
    int bad = 0;
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < n; t += (int64_t)gridDim.x * blockDim.x) {
        unsigned c = 0;
        const double a0 = m0[t], a1 = m1 ? m1[t] : 0.0, a2 = m2 ? m2[t] : 0.0;
        c |= a0 != 0.0 ? 1u : 0u;
        c |= a1 != 0.0 ? 2u : 0u;
        c |= a2 != 0.0 ? 4u : 0u;
        bad |= (a0 != 0.0 && a0 != 1.0) || (a1 != 0.0 && a1 != 1.0) || (a2 != 0.0 && a2 != 1.0);
        mcode[t] = (unsigned char)c;
        if (h2b) h2b[t] = h2[t] * bm1[t];
    }
    if (bad) *notbinary = 1;
}

struct HcgState {
    DevBuf<HcgScalars> sc;
    DevBuf<unsigned char> mcode;
    DevBuf<double> h2b, d, hist, partials;
    DevBuf<double> r[HCG_MAXR], p[HCG_MAXR], w[HCG_MAXR];
    DevBuf<int> flag;
};
inline HcgState &hcg_state()
{
    static HcgState s;
    return s;
}

inline int hcg_enabled()
{
    static int v = -1;
    if (v < 0) {
        const char *e = getenv("NEKB_HCG");
        v = e ? atoi(e) : 1;
    }
    return v;
}

template <int NRHS, int PIPES, int STAGES>
inline void hcg_launch_front(const HcgPtrs &P, const double *h1, const double *h2b, const double *d, int nel, HcgScalars *sc,
                             double *partials)
{
    Ctx &c = ctx();
    using L = HcgSmem<8, NRHS, PIPES, STAGES>;
    static bool configured = false;
    if (!configured) {
        NEKB_CUDA(cudaFuncSetAttribute(hcg_front_kernel<8, NRHS, PIPES, STAGES, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L::bytes));
        NEKB_CUDA(cudaFuncSetAttribute(hcg_front_kernel<8, NRHS, PIPES, STAGES, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L::bytes));
        configured = true;
    }
    const int grid = grid_for((nel + PIPES - 1) / PIPES, 1);
    if (h2b)
        hcg_front_kernel<8, NRHS, PIPES, STAGES, true><<<grid, 64 * NRHS * PIPES, L::bytes, c.stream>>>(P, c.g.p, h1, h2b, d, nel, sc, partials);
    else
        hcg_front_kernel<8, NRHS, PIPES, STAGES, false><<<grid, 64 * NRHS * PIPES, L::bytes, c.stream>>>(P, c.g.p, h1, h2b, d, nel, sc, partials);
    NEKB_LAUNCHED();
}

// Conditions of the fused path; anything else goes through cggo_run component by component.
inline bool hcg_applicable(int nrhs)
{
    Ctx &c = ctx();
    return hcg_enabled() && c.nx == 8 && (nrhs == 1 || nrhs == 3) && fdm_h1_kfldfdm() < 0;
}

// nrhs solves H x_c = f_c (Jacobi-PCG, the reference's recurrence per component).  x, f, mask: nrhs device pointers.
// tin[c]: the tolerance cggo would receive for component c.  niter_out[c] = niterhm of component c.  Returns false when the
// problem needs a branch the fused path does not provide (null-space correction, non-binary masks): nothing is modified then.
// hist_host (may be NULL): nrhs rows of 3*(min(maxit,900)+2) doubles: (rtz1, rbn2, rho) per executed iteration, as cggo_run.
inline bool hcg_run(int nrhs, double *const *x, const double *const *f, const double *h1, const double *h2, const double *const *mask,
                    const double *mult, const double *binv, int gs_handle, int nel, double vol, const double *tin, int maxit, int istep,
                    int *niter_out, double *hist_host)
{
    Ctx &c = ctx();
    HcgState &S = hcg_state();
    cudaStream_t s = c.stream;
    const int64_t n = (int64_t)nel * c.nxyz;
    const int maxcg = 900, niter = maxit < maxcg ? maxit : maxcg;
    // streaming kernels of this file: 4 resident CTAs per SM (<= 64 registers), one wave (round 1 launched 6 per SM of a
    // 90-register kernel: two resident, three waves, 50 % of the DRAM bandwidth)
    int grid = cg_grid(n);
    if (grid > c.num_sms * 4) grid = c.num_sms * 4;
    NEKB_REQUIRE(nrhs == 1 || nrhs == 3, "hcg: 1 or 3 right-hand sides");
    GsMap &h = gs_get(gs_handle);
    NEKB_REQUIRE(h.n == n, "hcg: gs handle was set up for a different vector length");
    auto misaligned = [](const void *q) { return (reinterpret_cast<uintptr_t>(q) & 15u) != 0; };  // 16-byte loads, TMA sources
    for (int k = 0; k < nrhs; k++)
        if (misaligned(x[k]) || misaligned(f[k])) return false;
    if (misaligned(h1) || misaligned(mult) || misaligned(binv)) return false;
    S.sc.ensure(1), S.mcode.ensure((size_t)n), S.d.ensure((size_t)n), S.flag.ensure(1);
    S.partials.ensure((size_t)8 * CG_PART_STRIDE);
    const int hstride = 3 * (niter + 2);
    S.hist.ensure((size_t)HCG_MAXR * hstride);
    c.partials.ensure(4 * CG_PART_STRIDE);
    CgScalars *csc = c.sc.p;

    // ifh2 (setfast :303-305) and the null-space test (:705-709) need max|h2| and min(mask)
    absmax_kernel<<<grid, CG_THREADS, 0, s>>>(h2, n, &csc->work[3], c.partials.p, &csc->counter[0]);
    NEKB_LAUNCHED();
    comm_allreduce_max(&csc->work[3], 1);
    double h2max = 0.0;
    NEKB_CUDA(cudaMemcpyAsync(&h2max, &csc->work[3], sizeof(double), cudaMemcpyDeviceToHost, s));
    NEKB_CUDA(cudaStreamSynchronize(s));
    if (h2max == 0.0) {
        for (int k = 0; k < nrhs; k++) {
            negmax_kernel<<<grid, CG_THREADS, 0, s>>>(mask[k], n, &csc->work[3], c.partials.p, &csc->counter[0]);
            NEKB_LAUNCHED();
            comm_allreduce_max(&csc->work[3], 1);
            double skmin = 0.0;
            NEKB_CUDA(cudaMemcpyAsync(&skmin, &csc->work[3], sizeof(double), cudaMemcpyDeviceToHost, s));
            NEKB_CUDA(cudaStreamSynchronize(s));
            if (-skmin > 0.0) return false;  // ifmcor: left to cggo_run
        }
    }
    const bool has_h2 = h2max > 0.0;
    // before hcg_prep_kernel reads bm1[t]: an unregistered bm1 must give this message, not a device fault
    if (has_h2) NEKB_REQUIRE(c.bm1.p != nullptr && c.bm1.n >= (size_t)n, "hcg: bm1 must be registered (h2 != 0)");
    if (has_h2) S.h2b.ensure((size_t)n);
    NEKB_CUDA(cudaMemsetAsync(S.flag.p, 0, sizeof(int), s));
    hcg_prep_kernel<<<grid, CG_THREADS, 0, s>>>(S.mcode.p, has_h2 ? S.h2b.p : nullptr, mask[0], nrhs > 1 ? mask[1] : nullptr,
                                                nrhs > 2 ? mask[2] : nullptr, h2, c.bm1.p, n, S.flag.p);
    NEKB_LAUNCHED();
    int notbinary = 0;
    NEKB_CUDA(cudaMemcpyAsync(&notbinary, S.flag.p, sizeof(int), cudaMemcpyDeviceToHost, s));
    NEKB_CUDA(cudaStreamSynchronize(s));
    if (notbinary) return false;

    setprec_run(S.d.p, h1, h2, nel, gs_handle);  // :690 (depends on h1, h2 only: shared by the components)

    // r = f, x = 0, p = 0 (:693-695); fmax = 0 -> that component returns at once with niterhm = 0 (:697-699)
    HcgScalars hs;
    memset(&hs, 0, sizeof hs);
    HcgPtrs P;
    memset(&P, 0, sizeof P);
    for (int k = 0; k < nrhs; k++) {
        S.r[k].ensure((size_t)n), S.p[k].ensure((size_t)n), S.w[k].ensure((size_t)n);
        P.x[k] = x[k], P.r[k] = S.r[k].p, P.p[k] = S.p[k].p, P.w[k] = S.w[k].p;
        NEKB_CUDA(cudaMemcpyAsync(S.r[k].p, f[k], sizeof(double) * (size_t)n, cudaMemcpyDeviceToDevice, s));
        NEKB_CUDA(cudaMemsetAsync(x[k], 0, sizeof(double) * (size_t)n, s));
        NEKB_CUDA(cudaMemsetAsync(S.p[k].p, 0, sizeof(double) * (size_t)n, s));
        absmax_kernel<<<grid, CG_THREADS, 0, s>>>(f[k], n, &csc->work[2], c.partials.p, &csc->counter[0]);
        NEKB_LAUNCHED();
        comm_allreduce_max(&csc->work[2], 1);
        double fmax = 0.0;
        NEKB_CUDA(cudaMemcpyAsync(&fmax, &csc->work[2], sizeof(double), cudaMemcpyDeviceToHost, s));
        NEKB_CUDA(cudaStreamSynchronize(s));
        hs.c[k].rtz1 = 1.0;  // :723
        if (fmax == 0.0) hs.c[k].done = 1, hs.c[k].niter = 0;
    }
    for (int k = nrhs; k < HCG_MAXR; k++) hs.c[k].done = 1;
    NEKB_CUDA(cudaMemcpyAsync(S.sc.p, &hs, sizeof hs, cudaMemcpyHostToDevice, s));
    NEKB_CUDA(cudaMemsetAsync(S.hist.p, 0, sizeof(double) * S.hist.n, s));
    HcgScalars *sc = S.sc.p;

    auto dots_allreduce = [&]() {
        if (c.nranks > 1)
            for (int k = 0; k < nrhs; k++) comm_allreduce_sum(&sc->c[k].work[0], 2);
    };
    auto update = [&](bool first) {
        if (nrhs == 3) {
            if (first)
                hcg_update_kernel<3, true><<<grid, CG_THREADS, 0, s>>>(P, S.mcode.p, mult, S.d.p, binv, n, sc, S.partials.p, S.hist.p, hstride);
            else
                hcg_update_kernel<3, false><<<grid, CG_THREADS, 0, s>>>(P, S.mcode.p, mult, S.d.p, binv, n, sc, S.partials.p, S.hist.p, hstride);
        } else {
            if (first)
                hcg_update_kernel<1, true><<<grid, CG_THREADS, 0, s>>>(P, S.mcode.p, mult, S.d.p, binv, n, sc, S.partials.p, S.hist.p, hstride);
            else
                hcg_update_kernel<1, false><<<grid, CG_THREADS, 0, s>>>(P, S.mcode.p, mult, S.d.p, binv, n, sc, S.partials.p, S.hist.p, hstride);
        }
        NEKB_LAUNCHED();
        dots_allreduce();
    };
    update(true);
    const char *rk = getenv("NEKB_HCG_RHO_KERNEL");
    const bool rho_kernel = rk && atoi(rk) != 0;
    int launched = 0;
    const int batch = 8;
    HcgScalars res;
    for (;;) {
        for (int b = 0; b < batch; b++) {
            hcg_check_kernel<<<1, 1, 0, s>>>(sc, nrhs, vol, tin[0], nrhs > 1 ? tin[1] : 0.0, nrhs > 2 ? tin[2] : 0.0, istep, niter,
                                             c.param[22], S.hist.p, hstride);
            NEKB_LAUNCHED();
            if (nrhs == 3)
                hcg_launch_front<3, 1, 3>(P, h1, has_h2 ? S.h2b.p : nullptr, S.d.p, nel, sc, S.partials.p);
            else
                hcg_launch_front<1, 2, 2>(P, h1, has_h2 ? S.h2b.p : nullptr, S.d.p, nel, sc, S.partials.p);
            for (int k = 0; k < nrhs; k++) gs_op(gs_handle, P.w[k], 1, nullptr);   // :797 dssum (done components: harmless)
            if (rho_kernel) {   // NEKB_HCG_RHO_KERNEL=1: rho from the assembled w in a pass of its own (round-1 form), for comparison
                if (nrhs == 3)
                    hcg_rho_kernel<3><<<grid, CG_THREADS, 0, s>>>(P, S.mcode.p, mult, n, sc, S.partials.p);
                else
                    hcg_rho_kernel<1><<<grid, CG_THREADS, 0, s>>>(P, S.mcode.p, mult, n, sc, S.partials.p);
                NEKB_LAUNCHED();
            }
            if (c.nranks > 1)
                for (int k = 0; k < nrhs; k++) comm_allreduce_sum(&sc->c[k].rho, 1);
            update(false);
            launched++;
        }
        NEKB_CUDA(cudaMemcpyAsync(&res, sc, sizeof res, cudaMemcpyDeviceToHost, s));
        NEKB_CUDA(cudaStreamSynchronize(s));
        if (res.alldone) break;
        NEKB_REQUIRE(launched <= niter + 3 * batch, "hcg: convergence flag never raised");
    }
    // the last batch may have run past the exit of every component: those launches found done = 1 and did nothing
    if (nrhs == 3)
        hcg_flush_kernel<3><<<grid, CG_THREADS, 0, s>>>(P, n, sc);
    else
        hcg_flush_kernel<1><<<grid, CG_THREADS, 0, s>>>(P, n, sc);
    NEKB_LAUNCHED();
    for (int k = 0; k < nrhs; k++) niter_out[k] = res.c[k].niter;
    if (hist_host) {
        std::vector<double> hh(S.hist.n);
        NEKB_CUDA(cudaMemcpyAsync(hh.data(), S.hist.p, sizeof(double) * hh.size(), cudaMemcpyDeviceToHost, s));
        NEKB_CUDA(cudaStreamSynchronize(s));
        for (int k = 0; k < nrhs; k++) memcpy(hist_host + (size_t)k * hstride, hh.data() + (size_t)k * hstride, sizeof(double) * hstride);
    } else
        NEKB_CUDA(cudaStreamSynchronize(s));
    return true;
}

// cggo with the fused path tried first (one right-hand side); falls back to the kernel-per-statement cggo_run for the
// branches hcg does not provide (Schwarz preconditioner, null-space correction, lx1 != 8, non-binary masks).
inline int cggo_solve_impl(const CggoArgs &a, double tin, int maxit, double *hist_host);
inline int cggo_solve(const CggoArgs &a, double tin, int maxit, double *hist_host)
{
    // the history always lands in ctx().last_hist as well (3 doubles per executed check: rtz1, rbn2, rho)
    Ctx &c = ctx();
    const int niter = maxit < 900 ? maxit : 900;
    c.last_hist.assign((size_t)3 * (niter + 2), 0.0);
    const int it = cggo_solve_impl(a, tin, maxit, c.last_hist.data());
    c.last_hist_rows = it + 1 <= niter + 1 ? it + 1 : niter + 1, c.last_hist_cols = 3;
    if (hist_host) memcpy(hist_host, c.last_hist.data(), sizeof(double) * 3 * (size_t)c.last_hist_rows);
    return it;
}
inline int cggo_solve_impl(const CggoArgs &a, double tin, int maxit, double *hist_host)
{
    tin = cggo_tin(tin);  // restol(ifield), hmholtz.f:676
    // :677 'PRES' with param(21) != 0: tol = |param(21)| (a negative tin still wins, :679)
    if (a.pres && tin >= 0.0 && ctx().param[21] != 0.0) tin = fabs(ctx().param[21]);
    if (hcg_applicable(1) && !a.pres) {
        double *xs[1] = {a.x};
        const double *fs[1] = {a.f}, *ms[1] = {a.mask};
        int it = 0;
        const int niter = maxit < 900 ? maxit : 900;
        std::vector<double> hh;
        if (hist_host) hh.resize((size_t)3 * (niter + 2));
        if (hcg_run(1, xs, fs, a.h1, a.h2, ms, a.mult, a.binv, a.gs_handle, a.nel, a.vol, &tin, maxit, a.istep, &it,
                    hist_host ? hh.data() : nullptr)) {
            if (hist_host) {
                const int rows = it + 1 <= niter + 1 ? it + 1 : niter + 1;
                for (int i = 0; i < 3 * rows; i++) hist_host[i] = hh[i];
            }
            return it;
        }
    }
    return cggo_run(a, tin, maxit, hist_host);
}

}  // namespace nekb
