// ax.cuh -- matrix-free Helmholtz / Poisson operator on (N+1)^3 element tiles.
//
// Replaces core/hmholtz.f:72-259 axhelm (general 3-D branch :191-217 + :225) and
// examples/bp5/bp5.usr:1278-1341 ax_e_bp5 / axhm1_bp5 (incl. the fused pap = sum p.Ap).
// The six mxm calls + 15 pointwise sweeps of the reference are fused into one pass:
//   ur,us,ut = D u ; (wr,ws,wt) = G (ur,us,ut) [* h1] ; w = D^T (wr,ws,wt) [+ h2 B u]
//
// Data layout in HBM: u, w as in Nek (i fastest, element slowest).  Geometric factors are stored
// element-major, component-next:  g[e][c][k][j][i], c = rr,rs,rt,ss,st,tt, so that each component
// of each k-slice is one coalesced 8*NX*NX-byte run and a whole element is one contiguous
// 6*NX^3*8-byte block (24 KB at N=7) for bulk (TMA) staging.
//
// Thread mapping (kernel v1): NX*NX threads per element, thread (i,j) owns the k-column u(i,j,:)
// in registers; the r/s contractions go through one shared-memory copy of the element tile, the t
// contraction stays in registers with D read from the constant bank.  EPB elements per CTA,
// persistent grid (multiple of the SM count) striding over the elements.
#pragma once
#include "ctx.cuh"

namespace nekb {

template <int NX, int EPB, bool HELM>
__global__ void __launch_bounds__(NX *NX *EPB)
    ax_kernel(const double *__restrict__ u, const double *__restrict__ g, double *__restrict__ w,
              const double *__restrict__ h1, const double *__restrict__ h2, const double *__restrict__ bm1, int nel,
              double *__restrict__ partials, unsigned *counter, double *pap_out)
{
    constexpr int N2 = NX * NX, N3 = NX * NX * NX;
    __shared__ double s_u[EPB][N3];
    __shared__ double s_wr[EPB][N3];  // r- and s-fluxes of all planes: one barrier per element instead of one per plane
    __shared__ double s_ws[EPB][N3];
    __shared__ double s_red[33];

    const int tid = threadIdx.x;
    const int es = tid / N2, ij = tid % N2, i = ij % NX, j = ij / NX;

    double Di[NX], Dj[NX], DTi[NX], DTj[NX];
#pragma unroll
    for (int m = 0; m < NX; m++) {
        Di[m] = c_D[i * NX + m];
        Dj[m] = c_D[j * NX + m];
        DTi[m] = c_D[m * NX + i];
        DTj[m] = c_D[m * NX + j];
    }

    double pap = 0.0;
    for (int e0 = blockIdx.x * EPB; e0 < nel; e0 += gridDim.x * EPB) {
        const int e = e0 + es;
        const bool act = e < nel;
        const size_t eo = (size_t)(act ? e : 0) * N3;
        const double *__restrict__ ge = g + 6 * eo;

        double ucol[NX], wcol[NX];
#pragma unroll
        for (int k = 0; k < NX; k++) {
            ucol[k] = act ? u[eo + k * N2 + ij] : 0.0;
            s_u[es][k * N2 + ij] = ucol[k];
            wcol[k] = 0.0;
        }
        __syncthreads();

#pragma unroll
        for (int k = 0; k < NX; k++) {
            const int q = k * N2 + ij;
            double G0 = 0, G1 = 0, G2 = 0, G3 = 0, G4 = 0, G5 = 0, hh = 1.0;
            if (act) {
                G0 = ge[0 * N3 + q];
                G1 = ge[1 * N3 + q];
                G2 = ge[2 * N3 + q];
                G3 = ge[3 * N3 + q];
                G4 = ge[4 * N3 + q];
                G5 = ge[5 * N3 + q];
                if (HELM) hh = h1[eo + q];
            }
            double ur = 0.0, us = 0.0, ut = 0.0;
#pragma unroll
            for (int m = 0; m < NX; m++) {
                ur = fma(Di[m], s_u[es][k * N2 + j * NX + m], ur);
                us = fma(Dj[m], s_u[es][k * N2 + m * NX + i], us);
                ut = fma(c_D[k * NX + m], ucol[m], ut);
            }
            double wr = fma(G0, ur, fma(G1, us, G2 * ut));
            double ws = fma(G1, ur, fma(G3, us, G4 * ut));
            double wt = fma(G2, ur, fma(G4, us, G5 * ut));
            if (HELM) {
                wr *= hh;
                ws *= hh;
                wt *= hh;
            }
            s_wr[es][q] = wr;
            s_ws[es][q] = ws;
#pragma unroll
            for (int m = 0; m < NX; m++) wcol[m] = fma(c_D[k * NX + m], wt, wcol[m]);
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < NX; k++) {
            double acc = wcol[k];
#pragma unroll
            for (int m = 0; m < NX; m++) {
                acc = fma(DTi[m], s_wr[es][k * N2 + j * NX + m], acc);
                acc = fma(DTj[m], s_ws[es][k * N2 + m * NX + i], acc);
            }
            wcol[k] = acc;
        }
        // no barrier needed here: the next element first rewrites s_u (not read above) and meets the barrier that follows
        // its load before any thread touches the flux tiles again
        if (act) {
#pragma unroll
            for (int k = 0; k < NX; k++) {
                const int q = k * N2 + ij;
                double v = wcol[k];
                if (HELM && h2 != nullptr) v = fma(h2[eo + q] * bm1[eo + q], ucol[k], v);
                w[eo + q] = v;
                pap = fma(ucol[k], v, pap);
            }
        }
    }
    if (pap_out != nullptr) {
        double b = block_reduce(pap, s_red);
        grid_reduce(b, partials, counter, s_red, [=](double t) { *pap_out = t; });
    }
}

// ---------------------------------------------------------------------------------------------- kernel v2 (TMA)
// Persistent CTAs of GROUPS x (NX*NX) threads.  Every group of NX*NX threads streams its own sequence of
// elements through a ring of STAGES shared-memory stages; a stage holds the element's six geometric-factor
// tiles and its u tile (one contiguous 6*NX^3*8-byte block + one NX^3*8-byte block in HBM), fetched with
// cp.async.bulk (TMA bulk copy, SASS UBLKCP) that signals an mbarrier with the byte count.  While a group
// computes on one stage the TMA engine fills the others, so HBM latency is hidden by the copy engine and not
// by occupancy; the factors never pass through registers until they are consumed.  Thread (i,j) owns the
// k-column of its element; r/s contractions read the u tile in place, the two D^T contractions exchange
// wr/ws through a double-buffered scratch slice with one 64-thread named barrier per k.
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "DONE:\n\t}" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void group_barrier(int id, int nthreads)
{
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

template <int NX, int GROUPS, int STAGES>
struct AxTmaSmem {
    static constexpr int N2 = NX * NX, N3 = NX * NX * NX;
    static constexpr size_t stage_doubles = 7 * (size_t)N3;                    // 6 factor tiles + u tile
    static constexpr size_t scratch_doubles = 4 * (size_t)N2;                  // wr[2], ws[2]
    static constexpr size_t group_doubles = STAGES * stage_doubles + scratch_doubles;
    static constexpr size_t bytes = GROUPS * group_doubles * sizeof(double) + GROUPS * STAGES * sizeof(uint64_t) + 64 * sizeof(double);
};

template <int NX, int GROUPS, int STAGES>
__global__ void __launch_bounds__(NX *NX *GROUPS, 1)
    ax_tma_kernel(const double *__restrict__ u, const double *__restrict__ g, double *__restrict__ w, int nel,
                  double *__restrict__ partials, unsigned *counter, double *pap_out)
{
    using L = AxTmaSmem<NX, GROUPS, STAGES>;
    constexpr int N2 = L::N2, N3 = L::N3;
    constexpr uint32_t G_BYTES = 6 * N3 * sizeof(double), U_BYTES = N3 * sizeof(double);
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double *smem = reinterpret_cast<double *>(smem_raw);
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + GROUPS * L::group_doubles);
    double *s_red = reinterpret_cast<double *>(bars + GROUPS * STAGES);

    const int grp = threadIdx.x / N2, ij = threadIdx.x % N2, i = ij % NX, j = ij / NX;
    double *gbase = smem + grp * L::group_doubles;
    double *s_w = gbase + STAGES * L::stage_doubles;  // [2][2][N2]: (parity, r|s)
    uint64_t *full = bars + grp * STAGES;
    const bool leader = ij == 0;

    if (threadIdx.x == 0) {
        for (int q = 0; q < GROUPS * STAGES; q++) mbar_init(&bars[q], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    double Di[NX], Dj[NX], DTi[NX], DTj[NX];
#pragma unroll
    for (int m = 0; m < NX; m++) {
        Di[m] = c_D[i * NX + m];
        Dj[m] = c_D[j * NX + m];
        DTi[m] = c_D[m * NX + i];
        DTj[m] = c_D[m * NX + j];
    }

    const int first = blockIdx.x * GROUPS + grp, stride = gridDim.x * GROUPS;
    auto issue = [&](int stage, int e) {
        double *st = gbase + stage * L::stage_doubles;
        mbar_expect_tx(&full[stage], G_BYTES + U_BYTES);
        bulk_g2s(st, g + (size_t)e * 6 * N3, G_BYTES, &full[stage]);
        bulk_g2s(st + 6 * N3, u + (size_t)e * N3, U_BYTES, &full[stage]);
    };
    if (leader) {
#pragma unroll
        for (int s = 0; s < STAGES; s++) {
            const int e = first + s * stride;
            if (e < nel) issue(s, e);
        }
    }

    double pap = 0.0;
    int it = 0;
    for (int e = first; e < nel; e += stride, it++) {
        const int stage = it % STAGES;
        const uint32_t parity = (uint32_t)(it / STAGES) & 1u;
        mbar_wait(&full[stage], parity);
        double *sgw = gbase + stage * L::stage_doubles;
        const double *sg = sgw;
        const double *su = sg + 6 * N3;

        double ucol[NX], wcol[NX];
#pragma unroll
        for (int k = 0; k < NX; k++) {
            ucol[k] = su[k * N2 + ij];
            wcol[k] = 0.0;
        }
#pragma unroll
        for (int k = 0; k < NX; k++) {
            const int q = k * N2 + ij;
            const double G0 = sg[0 * N3 + q], G1 = sg[1 * N3 + q], G2 = sg[2 * N3 + q], G3 = sg[3 * N3 + q],
                         G4 = sg[4 * N3 + q], G5 = sg[5 * N3 + q];
            double ur = 0.0, us = 0.0, ut = 0.0;
#pragma unroll
            for (int m = 0; m < NX; m++) {
                ur = fma(Di[m], su[k * N2 + j * NX + m], ur);
                us = fma(Dj[m], su[k * N2 + m * NX + i], us);
                ut = fma(c_D[k * NX + m], ucol[m], ut);
            }
            const double wr = fma(G0, ur, fma(G1, us, G2 * ut));
            const double ws = fma(G1, ur, fma(G3, us, G4 * ut));
            const double wt = fma(G2, ur, fma(G4, us, G5 * ut));
            // the factor entries at q are read by this thread only: their slots take the r- and s-flux of the plane, so
            // one group barrier per element (instead of one per plane) separates the fluxes from their D^T contractions
            sgw[0 * N3 + q] = wr;
            sgw[1 * N3 + q] = ws;
#pragma unroll
            for (int m = 0; m < NX; m++) wcol[m] = fma(c_D[k * NX + m], wt, wcol[m]);
        }
        group_barrier(1 + grp, N2);
#pragma unroll
        for (int k = 0; k < NX; k++) {
            double acc = wcol[k];
#pragma unroll
            for (int m = 0; m < NX; m++) {
                acc = fma(DTi[m], sgw[k * N2 + j * NX + m], acc);
                acc = fma(DTj[m], sgw[N3 + k * N2 + m * NX + i], acc);
            }
            wcol[k] = acc;
        }
        double *__restrict__ we = w + (size_t)e * N3;
#pragma unroll
        for (int k = 0; k < NX; k++) {
            we[k * N2 + ij] = wcol[k];
            pap = fma(ucol[k], wcol[k], pap);
        }
        // every thread of the group is past its last read of this stage (and of the scratch slices) once it
        // has passed this barrier; only then may the TMA engine overwrite the stage
        group_barrier(1 + grp, N2);
        if (leader) {
            const int en = e + STAGES * stride;
            if (en < nel) {
                // flux values were written into the stage through the generic proxy: order them before the refill
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                issue(stage, en);
            }
        }
    }
    if (pap_out != nullptr) {
        double b = block_reduce(pap, s_red);
        grid_reduce(b, partials, counter, s_red, [=](double t) { *pap_out = t; });
    }
}

// ---------------------------------------------------------------------------------------------- kernel v3 (TMA, CG-fused)
// The front half of a cggos iteration in one pass over the element tiles (examples/bp5/bp5.usr:855-856 of the
// PREVIOUS iteration, :879-884, then :848 of this one):
//     u += alpha * p_old ;  p = r + beta * p_old ;  w = A p ;  pap = sum p.w
// Compared with separate kernels this removes one read of p and the read-modify-write of u from the vector
// update.  FIRST (iteration 1): p = r, u untouched.  alpha and beta come from the device-resident CG scalars.
template <int NX, int GROUPS, int STAGES>
struct AxCgSmem {
    static constexpr int N2 = NX * NX, N3 = NX * NX * NX;
    static constexpr size_t stage_doubles = 9 * (size_t)N3;                    // 6 factor tiles + r, p, u tiles
    static constexpr size_t scratch_doubles = 4 * (size_t)N2;
    static constexpr size_t group_doubles = STAGES * stage_doubles + scratch_doubles;
    static constexpr size_t bytes = GROUPS * group_doubles * sizeof(double) + GROUPS * STAGES * sizeof(uint64_t) + 64 * sizeof(double);
};

template <int NX, int GROUPS, int STAGES, bool FIRST>
__global__ void __launch_bounds__(NX *NX *GROUPS, 1)
    ax_cg_kernel(const double *__restrict__ r, double *__restrict__ p, double *__restrict__ u,
                 const double *__restrict__ g, double *__restrict__ w, int nel, const CgScalars *__restrict__ sc,
                 double *__restrict__ partials, unsigned *counter, double *pap_out)
{
    using L = AxCgSmem<NX, GROUPS, STAGES>;
    constexpr int N2 = L::N2, N3 = L::N3;
    constexpr uint32_t G_BYTES = 6 * N3 * sizeof(double), T_BYTES = N3 * sizeof(double);
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double *smem = reinterpret_cast<double *>(smem_raw);
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + GROUPS * L::group_doubles);
    double *s_red = reinterpret_cast<double *>(bars + GROUPS * STAGES);

    const int grp = threadIdx.x / N2, ij = threadIdx.x % N2, i = ij % NX, j = ij / NX;
    double *gbase = smem + grp * L::group_doubles;
    double *s_w = gbase + STAGES * L::stage_doubles;
    uint64_t *full = bars + grp * STAGES;
    const bool leader = ij == 0;

    if (threadIdx.x == 0) {
        for (int q = 0; q < GROUPS * STAGES; q++) mbar_init(&bars[q], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    // bp5.usr:853 alpha = rpp1/pap of the previous iteration; :879-881 beta = rpp1/rpp2
    const double alpha = FIRST ? 0.0 : sc->alpha;
    const double beta = FIRST ? 0.0 : sc->work[1] / sc->rtz1;

    double Di[NX], Dj[NX], DTi[NX], DTj[NX];
#pragma unroll
    for (int m = 0; m < NX; m++) {
        Di[m] = c_D[i * NX + m];
        Dj[m] = c_D[j * NX + m];
        DTi[m] = c_D[m * NX + i];
        DTj[m] = c_D[m * NX + j];
    }

    const int first = blockIdx.x * GROUPS + grp, stride = gridDim.x * GROUPS;
    auto issue = [&](int stage, int e) {
        double *st = gbase + stage * L::stage_doubles;
        mbar_expect_tx(&full[stage], G_BYTES + (FIRST ? 1 : 3) * T_BYTES);
        bulk_g2s(st, g + (size_t)e * 6 * N3, G_BYTES, &full[stage]);
        bulk_g2s(st + 6 * N3, r + (size_t)e * N3, T_BYTES, &full[stage]);
        if (!FIRST) {
            bulk_g2s(st + 7 * N3, p + (size_t)e * N3, T_BYTES, &full[stage]);
            bulk_g2s(st + 8 * N3, u + (size_t)e * N3, T_BYTES, &full[stage]);
        }
    };
    if (leader) {
#pragma unroll
        for (int s = 0; s < STAGES; s++) {
            const int e = first + s * stride;
            if (e < nel) issue(s, e);
        }
    }

    double pap = 0.0;
    int it = 0;
    for (int e = first; e < nel; e += stride, it++) {
        const int stage = it % STAGES;
        const uint32_t parity = (uint32_t)(it / STAGES) & 1u;
        mbar_wait(&full[stage], parity);
        double *sg = gbase + stage * L::stage_doubles;
        double *sr = sg + 6 * N3, *sp = sg + 7 * N3;
        const double *su = sg + 8 * N3;
        const double *sf = FIRST ? sr : sp;  // the tile holding the search direction p of this iteration

        double ucol[NX], wcol[NX];
        double *__restrict__ pe = p + (size_t)e * N3;
        if (FIRST) {
#pragma unroll
            for (int k = 0; k < NX; k++) {
                ucol[k] = sr[k * N2 + ij];
                pe[k * N2 + ij] = ucol[k];
                wcol[k] = 0.0;
            }
        } else {
            double *__restrict__ ue = u + (size_t)e * N3;
#pragma unroll
            for (int k = 0; k < NX; k++) {
                const double po = sp[k * N2 + ij];
                const double pn = fma(beta, po, sr[k * N2 + ij]);
                ue[k * N2 + ij] = fma(alpha, po, su[k * N2 + ij]);
                sp[k * N2 + ij] = pn;
                pe[k * N2 + ij] = pn;
                ucol[k] = pn;
                wcol[k] = 0.0;
            }
            group_barrier(1 + grp, N2);  // the p tile is complete
        }
        // The r and u tiles of this stage are dead once p is built (FIRST: the unused p and u slots): they take the
        // r- and s-fluxes of all NX planes, so the two D^T contractions need ONE group barrier per element instead of one
        // per plane.
        double *swr = FIRST ? sg + 7 * N3 : sr, *sws = sg + 8 * N3;
#pragma unroll
        for (int k = 0; k < NX; k++) {
            const int q = k * N2 + ij;
            const double G0 = sg[0 * N3 + q], G1 = sg[1 * N3 + q], G2 = sg[2 * N3 + q], G3 = sg[3 * N3 + q],
                         G4 = sg[4 * N3 + q], G5 = sg[5 * N3 + q];
            double ur = 0.0, us = 0.0, ut = 0.0;
#pragma unroll
            for (int m = 0; m < NX; m++) {
                ur = fma(Di[m], sf[k * N2 + j * NX + m], ur);
                us = fma(Dj[m], sf[k * N2 + m * NX + i], us);
                ut = fma(c_D[k * NX + m], ucol[m], ut);
            }
            const double wr = fma(G0, ur, fma(G1, us, G2 * ut));
            const double ws = fma(G1, ur, fma(G3, us, G4 * ut));
            const double wt = fma(G2, ur, fma(G4, us, G5 * ut));
            swr[q] = wr;
            sws[q] = ws;
#pragma unroll
            for (int m = 0; m < NX; m++) wcol[m] = fma(c_D[k * NX + m], wt, wcol[m]);
        }
        group_barrier(1 + grp, N2);
#pragma unroll
        for (int k = 0; k < NX; k++) {
            double acc = wcol[k];
#pragma unroll
            for (int m = 0; m < NX; m++) {
                acc = fma(DTi[m], swr[k * N2 + j * NX + m], acc);
                acc = fma(DTj[m], sws[k * N2 + m * NX + i], acc);
            }
            wcol[k] = acc;
        }
        double *__restrict__ we = w + (size_t)e * N3;
#pragma unroll
        for (int k = 0; k < NX; k++) {
            we[k * N2 + ij] = wcol[k];
            pap = fma(ucol[k], wcol[k], pap);
        }
        group_barrier(1 + grp, N2);
        if (leader) {
            const int en = e + STAGES * stride;
            if (en < nel) {
                // the p tile was written through the generic proxy: order those writes before the async-proxy refill
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                issue(stage, en);
            }
        }
    }
    if (pap_out != nullptr) {
        double b = block_reduce(pap, s_red);
        grid_reduce(b, partials, counter, s_red, [=](double t) { *pap_out = t; });
    }
}

// ---------------------------------------------------------------------------------------------- kernel v4 (affine elements)
// On an affine element (a parallelepiped: every element of a genbox mesh, uniform or stretched) the Jacobian matrix is
// constant, so each of the six factors is one number per element times the quadrature weight w_i w_j w_k of the node
// (core/coef.f:633-784: G = (cofactor products) / J * w3m1 with constant cofactors and J).  The factors of such elements
// need not be streamed per node: the kernel below is ax_cg_kernel with the 24 KB factor block of a stage replaced by one
// 64-byte record of six constants (HBM traffic of the kernel: 6 instead of 12 words per node).  Whether a mesh qualifies is
// DECIDED FROM THE REGISTERED FACTORS THEMSELVES (ax_affine_ensure: every node of every element within 1e-13 of
// constant * w3), never assumed; anything else keeps the general kernel.
constexpr int AFF_REC = 8;   // doubles per element record (6 used)

template <int NX>
__global__ void __launch_bounds__(NX *NX) ax_affine_check_kernel(const double *__restrict__ g, double *__restrict__ gc, int nel,
                                                                unsigned long long *maxdev_bits)
{
    constexpr int N2 = NX * NX, N3 = NX * NX * NX;
    __shared__ double s_part[6][N2];
    __shared__ double s_ref[6];
    __shared__ double s_max;
    const int e = blockIdx.x, ij = threadIdx.x, i = ij % NX, j = ij / NX;
    const double *ge = g + (size_t)e * 6 * N3;
    // constants = mean over the tile of G_c / w3 (the registered factors carry the rounding noise of the numerical
    // differentiation they come from -- up to 2e-12 relative on a 64^3 box; the mean halves what one node would carry)
    for (int c = 0; c < 6; c++) {
        double a = 0.0;
        for (int k = 0; k < NX; k++) a += ge[c * N3 + k * N2 + ij] / (c_w[i] * c_w[j] * c_w[k]);
        s_part[c][ij] = a;
    }
    __syncthreads();
    if (ij < 6) {
        double a = 0.0;
        for (int t = 0; t < N2; t++) a += s_part[ij][t];
        s_ref[ij] = a / (double)N3;
    }
    __syncthreads();
    if (ij == 0) {
        double m = 0.0;
        for (int c = 0; c < 6; c++) m = fmax(m, fabs(s_ref[c]));
        s_max = m;
        for (int c = 0; c < 6; c++) gc[(size_t)e * AFF_REC + c] = s_ref[c];
        gc[(size_t)e * AFF_REC + 6] = gc[(size_t)e * AFF_REC + 7] = 0.0;
    }
    __syncthreads();
    double dev = s_max > 0.0 ? 0.0 : 1.0;
    for (int k = 0; k < NX; k++) {
        const double w3 = c_w[i] * c_w[j] * c_w[k];
        for (int c = 0; c < 6; c++) {
            const double d = fabs(ge[c * N3 + k * N2 + ij] - s_ref[c] * w3) / (s_max * w3);
            dev = (d > dev || d != d) ? (d != d ? 1.0 : d) : dev;
        }
    }
    atomicMax(maxdev_bits, (unsigned long long)__double_as_longlong(dev));   // non-negative doubles order like their bits
}

template <int NX, int GROUPS, int STAGES>
struct AxCgAffSmem {
    static constexpr int N2 = NX * NX, N3 = NX * NX * NX;
    static constexpr size_t stage_doubles = 3 * (size_t)N3 + AFF_REC;          // r, p, u tiles + the element's constants
    static constexpr size_t group_doubles = STAGES * stage_doubles;
    static constexpr size_t bytes = GROUPS * group_doubles * sizeof(double) + GROUPS * STAGES * sizeof(uint64_t) + 64 * sizeof(double);
};

template <int NX, int GROUPS, int STAGES, bool FIRST>
__global__ void __launch_bounds__(NX *NX *GROUPS, 1)
    ax_cg_affine_kernel(const double *__restrict__ r, double *__restrict__ p, double *__restrict__ u,
                        const double *__restrict__ gc, double *__restrict__ w, int nel, const CgScalars *__restrict__ sc,
                        double *__restrict__ partials, unsigned *counter, double *pap_out)
{
    using L = AxCgAffSmem<NX, GROUPS, STAGES>;
    constexpr int N2 = L::N2, N3 = L::N3;
    constexpr uint32_t T_BYTES = N3 * sizeof(double), C_BYTES = AFF_REC * sizeof(double);
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double *smem = reinterpret_cast<double *>(smem_raw);
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + GROUPS * L::group_doubles);
    double *s_red = reinterpret_cast<double *>(bars + GROUPS * STAGES);

    const int grp = threadIdx.x / N2, ij = threadIdx.x % N2, i = ij % NX, j = ij / NX;
    double *gbase = smem + grp * L::group_doubles;
    uint64_t *full = bars + grp * STAGES;
    const bool leader = ij == 0;

    if (threadIdx.x == 0) {
        for (int q = 0; q < GROUPS * STAGES; q++) mbar_init(&bars[q], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    const double alpha = FIRST ? 0.0 : sc->alpha;
    const double beta = FIRST ? 0.0 : sc->work[1] / sc->rtz1;
    const double wij = c_w[i] * c_w[j];

    double Di[NX], Dj[NX], DTi[NX], DTj[NX];
#pragma unroll
    for (int m = 0; m < NX; m++) {
        Di[m] = c_D[i * NX + m];
        Dj[m] = c_D[j * NX + m];
        DTi[m] = c_D[m * NX + i];
        DTj[m] = c_D[m * NX + j];
    }

    const int first = blockIdx.x * GROUPS + grp, stride = gridDim.x * GROUPS;
    auto issue = [&](int stage, int e) {
        double *st = gbase + stage * L::stage_doubles;
        mbar_expect_tx(&full[stage], C_BYTES + (FIRST ? 1 : 3) * T_BYTES);
        bulk_g2s(st + 3 * N3, gc + (size_t)e * AFF_REC, C_BYTES, &full[stage]);
        bulk_g2s(st, r + (size_t)e * N3, T_BYTES, &full[stage]);
        if (!FIRST) {
            bulk_g2s(st + N3, p + (size_t)e * N3, T_BYTES, &full[stage]);
            bulk_g2s(st + 2 * N3, u + (size_t)e * N3, T_BYTES, &full[stage]);
        }
    };
    if (leader) {
#pragma unroll
        for (int s = 0; s < STAGES; s++) {
            const int e = first + s * stride;
            if (e < nel) issue(s, e);
        }
    }

    double pap = 0.0;
    int it = 0;
    for (int e = first; e < nel; e += stride, it++) {
        const int stage = it % STAGES;
        const uint32_t parity = (uint32_t)(it / STAGES) & 1u;
        mbar_wait(&full[stage], parity);
        double *sr = gbase + stage * L::stage_doubles, *sp = sr + N3, *su = sr + 2 * N3;
        const double *sgc = sr + 3 * N3;
        const double *sf = FIRST ? sr : sp;  // the tile holding the search direction p of this iteration
        // the six constants of the element, scaled by this thread's w_i w_j (w_k follows per plane)
        const double c0 = sgc[0] * wij, c1 = sgc[1] * wij, c2 = sgc[2] * wij, c3 = sgc[3] * wij, c4 = sgc[4] * wij, c5 = sgc[5] * wij;

        double ucol[NX], wcol[NX];
        double *__restrict__ pe = p + (size_t)e * N3;
        if (FIRST) {
#pragma unroll
            for (int k = 0; k < NX; k++) {
                ucol[k] = sr[k * N2 + ij];
                pe[k * N2 + ij] = ucol[k];
                wcol[k] = 0.0;
            }
        } else {
            double *__restrict__ ue = u + (size_t)e * N3;
#pragma unroll
            for (int k = 0; k < NX; k++) {
                const double po = sp[k * N2 + ij];
                const double pn = fma(beta, po, sr[k * N2 + ij]);
                ue[k * N2 + ij] = fma(alpha, po, su[k * N2 + ij]);
                sp[k * N2 + ij] = pn;
                pe[k * N2 + ij] = pn;
                ucol[k] = pn;
                wcol[k] = 0.0;
            }
            group_barrier(1 + grp, N2);  // the p tile is complete
        }
        // r and u tiles are dead once p is built (FIRST: the unused p and u slots): they take the r- and s-fluxes
        double *swr = FIRST ? sp : sr, *sws = su;
#pragma unroll
        for (int k = 0; k < NX; k++) {
            const int q = k * N2 + ij;
            const double wk = c_w[k];
            double ur = 0.0, us = 0.0, ut = 0.0;
#pragma unroll
            for (int m = 0; m < NX; m++) {
                ur = fma(Di[m], sf[k * N2 + j * NX + m], ur);
                us = fma(Dj[m], sf[k * N2 + m * NX + i], us);
                ut = fma(c_D[k * NX + m], ucol[m], ut);
            }
            // G_ab(i,j,k) = c_ab * w_i w_j w_k: the same three products as the general kernel, with the common weight last
            const double wr = fma(c0, ur, fma(c1, us, c2 * ut)) * wk;
            const double ws = fma(c1, ur, fma(c3, us, c4 * ut)) * wk;
            const double wt = fma(c2, ur, fma(c4, us, c5 * ut)) * wk;
            swr[q] = wr;
            sws[q] = ws;
#pragma unroll
            for (int m = 0; m < NX; m++) wcol[m] = fma(c_D[k * NX + m], wt, wcol[m]);
        }
        group_barrier(1 + grp, N2);
#pragma unroll
        for (int k = 0; k < NX; k++) {
            double acc = wcol[k];
#pragma unroll
            for (int m = 0; m < NX; m++) {
                acc = fma(DTi[m], swr[k * N2 + j * NX + m], acc);
                acc = fma(DTj[m], sws[k * N2 + m * NX + i], acc);
            }
            wcol[k] = acc;
        }
        double *__restrict__ we = w + (size_t)e * N3;
#pragma unroll
        for (int k = 0; k < NX; k++) {
            we[k * N2 + ij] = wcol[k];
            pap = fma(ucol[k], wcol[k], pap);
        }
        group_barrier(1 + grp, N2);
        if (leader) {
            const int en = e + STAGES * stride;
            if (en < nel) {
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                issue(stage, en);
            }
        }
    }
    if (pap_out != nullptr) {
        double b = block_reduce(pap, s_red);
        grid_reduce(b, partials, counter, s_red, [=](double t) { *pap_out = t; });
    }
}

// Decides (once per registered geometry) whether every element is affine and, if so, extracts the per-element constants.
inline bool ax_affine_ensure()
{
    Ctx &c = ctx();
    const char *env = getenv("NEKB_AX_AFFINE");      // read per solve: 0 keeps the general (per-node factors) kernel
    if (env && atoi(env) == 0) return false;
    if (c.affine_gen == c.geom_gen) return c.affine;
    c.affine_gen = c.geom_gen;
    c.affine = false;
    if (c.nx != 8 || !c.have_geom || !c.have_gll || c.nelt < 1) return false;
    NEKB_CUDA(cudaMemcpyToSymbolAsync(c_w, c.w_host.data(), sizeof(double) * c.nx, 0, cudaMemcpyHostToDevice, c.stream));
    c.gc.alloc((size_t)c.nelt * AFF_REC);
    DevBuf<unsigned long long> md;
    md.alloc(1);
    md.zero(c.stream);
    ax_affine_check_kernel<8><<<c.nelt, 64, 0, c.stream>>>(c.g.p, c.gc.p, c.nelt, md.p);
    NEKB_LAUNCHED();
    unsigned long long bits = 0;
    md.download(&bits, 1, c.stream);
    memcpy(&c.affine_maxdev, &bits, sizeof(double));
    // Tolerance: a deviation d of the factors moves A u by about 0.3 d (measured); the parity budget of one
    // application is 1e-12, so the affine kernel is used only while d <= 3e-12 (NEKB_AX_AFFINE_TOL overrides).
    const char *te = getenv("NEKB_AX_AFFINE_TOL");
    const double tol = te ? atof(te) : 3e-12;
    c.affine = c.affine_maxdev <= tol;
    if (!c.affine) c.gc.release();
    return c.affine;
}

template <int NX, int GROUPS, int STAGES>
inline void launch_ax_cg_affine(const double *r, double *p, double *u, double *w, int nel, bool first, double *pap_out)
{
    Ctx &c = ctx();
    using L = AxCgAffSmem<NX, GROUPS, STAGES>;
    static bool configured = false;
    if (!configured) {
        NEKB_CUDA(cudaFuncSetAttribute(ax_cg_affine_kernel<NX, GROUPS, STAGES, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L::bytes));
        NEKB_CUDA(cudaFuncSetAttribute(ax_cg_affine_kernel<NX, GROUPS, STAGES, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L::bytes));
        configured = true;
    }
    const int grid = grid_for((nel + GROUPS - 1) / GROUPS, 1);
    if (first)
        ax_cg_affine_kernel<NX, GROUPS, STAGES, true><<<grid, NX * NX * GROUPS, L::bytes, c.stream>>>(
            r, p, u, c.gc.p, w, nel, c.sc.p, c.partials.p, &c.sc.p->counter[0], pap_out);
    else
        ax_cg_affine_kernel<NX, GROUPS, STAGES, false><<<grid, NX * NX * GROUPS, L::bytes, c.stream>>>(
            r, p, u, c.gc.p, w, nel, c.sc.p, c.partials.p, &c.sc.p->counter[0], pap_out);
    NEKB_LAUNCHED();
}

// ---------------------------------------------------------------------------------------------- kernel v5 (affine, warp per element, DMMA)
// ncu of kernel v4 (profiles/r2o_ncu_summary.md): DRAM 59 % active, the shared-memory pipe and the FP64 pipe busy in turn
// (858 + 514 of 1466 cycles per element and SM, profiles/r2p_ubench_dmma.txt: LDS.64 = 2 clk, LDS.128 = 1 clk per warp
// instruction), two warps per scheduler: the contraction, not the bandwidth, bounds it -- the case for which the north star
// allows FP64 tensor-core instructions.  Here ONE WARP owns an element and the two in-plane contractions of every k-plane
// (D u and u D^T, and their transposes on the way back) are mma.sync.m8n8k4.f64 (SASS DMMA.8x8x4; same flop rate as DFMA on
// B200, but a fragment is ONE shared-memory word per lane instead of eight):
//   lane = (g = lane / 4, t = lane % 4) owns the two k-columns (i = g, j = t) and (i = g, j = t + 4): the C-fragment layout
//   (row g, columns 2t, 2t+1) with matrix column c standing for j = pi(c), pi(2q + s) = q + 4s.
//   ur(i,j) = sum_m D(i,m) p(m,j,k):  A = D (registers), B = p[k][pi(g)][2t .. 2t+1]: ONE LDS.128 per plane (contraction index
//             m = 2t + s in step s, so a lane's two B elements are adjacent; the 32 lanes read the plane's 512 bytes once)
//   us(i,j) = sum_m p(i,m,k) D(j,m):  A(row g, col t, step s) = p(i = g, j = t + 4s, k) = the lane's OWN node: no load at all
//   ut        from the lane's own k-columns in registers with D(k,m) from the constant bank, as in v4
//   back:  D^T wr through shared memory (own-address STS.64, one LDS.128 fragment per plane), D^T ws from the lane's own values,
//          D^T wt accumulated in registers; the DMMA accumulator starts from that register sum.
// Per element a lane issues 112 LDS/STS.64 + 16 LDS.128 (v4: 2 x (154 + 67 + 27)), 64 DMMA and ~520 DFMA; no block or group
// barrier (only __syncwarp).  The own-node accesses (words t*8 + g of a plane) are two-way bank conflicts per half warp (ncu:
// 95 M conflicts per launch); the alternative ownership (j = g, i = 2t, 2t+1: own nodes one 16-byte word, conflict-free, but two
// LDS.64 per B fragment) measured 2-4 % SLOWER (profiles/r2z_*: 1.12-1.14 vs 1.07-1.09 ms) and was dropped -- the kernel is
// DRAM-bound, not shared-memory-bound.
__device__ __forceinline__ void dmma884(double &c0, double &c1, double a, double b)
{
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

template <int WARPS, int STAGES>
struct AxCgAffMmaSmem {
    static constexpr int N3 = 512;
    static constexpr size_t stage_doubles = 3 * (size_t)N3 + AFF_REC;          // r, p, u tiles + the element's constants
    static constexpr size_t warp_doubles = STAGES * stage_doubles;
    static constexpr size_t bytes = WARPS * warp_doubles * sizeof(double) + WARPS * STAGES * sizeof(uint64_t) + 64 * sizeof(double);
};

template <int WARPS, int STAGES, bool FIRST>
__global__ void __launch_bounds__(32 * WARPS, 1)
    ax_cg_affine_mma_kernel(const double *__restrict__ r, double *__restrict__ p, double *__restrict__ u,
                            const double *__restrict__ gc, double *__restrict__ w, int nel,
                            const CgScalars *__restrict__ sc, double *__restrict__ partials, unsigned *counter, double *pap_out)
{
    using L = AxCgAffMmaSmem<WARPS, STAGES>;
    constexpr int NX = 8, N2 = 64, N3 = 512;
    constexpr uint32_t T_BYTES = N3 * sizeof(double), C_BYTES = AFF_REC * sizeof(double);
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double *smem = reinterpret_cast<double *>(smem_raw);
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + WARPS * L::warp_doubles);
    double *s_red = reinterpret_cast<double *>(bars + WARPS * STAGES);

    const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    const int pg = (g >> 1) + 4 * (g & 1);   // pi(g): the j a B-fragment column / C-fragment column g stands for
    double *wbase = smem + wid * L::warp_doubles;
    uint64_t *full = bars + wid * STAGES;

    if (threadIdx.x == 0) {
        for (int q = 0; q < WARPS * STAGES; q++) mbar_init(&bars[q], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    const double alpha = FIRST ? 0.0 : sc->alpha;
    const double beta = FIRST ? 0.0 : sc->work[1] / sc->rtz1;
    // operand fragments of D (step s of the two k = 4 steps of an 8-long contraction)
    double dA[2], dB[2], dAt[2], dBt[2], wij[2];
#pragma unroll
    for (int s = 0; s < 2; s++) {
        dA[s] = c_D[g * NX + 2 * t + s];           // ur : A(i = g, m = 2t + s) = D(i, m)
        dB[s] = c_D[pg * NX + t + 4 * s];          // us : B(m = t + 4s, j = pi(g)) = D(j, m)
        dAt[s] = c_D[(2 * t + s) * NX + g];        // D^T wr : A(i = g, m = 2t + s) = D(m, i)
        dBt[s] = c_D[(t + 4 * s) * NX + pg];       // D^T ws : B(m = t + 4s, j = pi(g)) = D(m, j)
        wij[s] = c_w[g] * c_w[t + 4 * s];
    }
    const int own0 = t * NX + g, own1 = (t + 4) * NX + g;   // the lane's two nodes in a plane: 32 consecutive words per warp
    const int frag = pg * NX + 2 * t;                      // B fragment (two adjacent words) in a plane

    const int first = blockIdx.x * WARPS + wid, stride = gridDim.x * WARPS;
    auto issue = [&](int stage, int e) {
        double *st = wbase + stage * L::stage_doubles;
        mbar_expect_tx(&full[stage], C_BYTES + (FIRST ? 1 : 3) * T_BYTES);
        bulk_g2s(st + 3 * N3, gc + (size_t)e * AFF_REC, C_BYTES, &full[stage]);
        bulk_g2s(st, r + (size_t)e * N3, T_BYTES, &full[stage]);
        if (!FIRST) {
            bulk_g2s(st + N3, p + (size_t)e * N3, T_BYTES, &full[stage]);
            bulk_g2s(st + 2 * N3, u + (size_t)e * N3, T_BYTES, &full[stage]);
        }
    };
    if (lane == 0) {
#pragma unroll
        for (int s = 0; s < STAGES; s++) {
            const int e = first + s * stride;
            if (e < nel) issue(s, e);
        }
    }

    double pap = 0.0;
    int it = 0;
    for (int e = first; e < nel; e += stride, it++) {
        const int stage = it % STAGES;
        const uint32_t parity = (uint32_t)(it / STAGES) & 1u;
        mbar_wait(&full[stage], parity);
        double *sr = wbase + stage * L::stage_doubles, *sp = sr + N3, *su = sr + 2 * N3;
        const double *sgc = sr + 3 * N3;
        const double *sf = FIRST ? sr : sp;    // the tile holding the search direction p of this iteration
        double *swr = FIRST ? sp : sr;         // dead tile that takes the r-fluxes
        const double c0 = sgc[0], c1 = sgc[1], c2 = sgc[2], c3 = sgc[3], c4 = sgc[4], c5 = sgc[5];

        // ---- phase 0: u += alpha p_old ; p = r + beta p_old at the lane's own nodes
        double pc[2][NX];
        double *__restrict__ pe = p + (size_t)e * N3;
        if (FIRST) {
#pragma unroll
            for (int k = 0; k < NX; k++) {
                pc[0][k] = sr[k * N2 + own0];
                pc[1][k] = sr[k * N2 + own1];
                pe[k * N2 + own0] = pc[0][k];
                pe[k * N2 + own1] = pc[1][k];
            }
        } else {
            double *__restrict__ ue = u + (size_t)e * N3;
#pragma unroll
            for (int k = 0; k < NX; k++) {
                const double po0 = sp[k * N2 + own0], po1 = sp[k * N2 + own1];
                const double pn0 = fma(beta, po0, sr[k * N2 + own0]), pn1 = fma(beta, po1, sr[k * N2 + own1]);
                ue[k * N2 + own0] = fma(alpha, po0, su[k * N2 + own0]);
                ue[k * N2 + own1] = fma(alpha, po1, su[k * N2 + own1]);
                sp[k * N2 + own0] = pn0;
                sp[k * N2 + own1] = pn1;
                pe[k * N2 + own0] = pn0;
                pe[k * N2 + own1] = pn1;
                pc[0][k] = pn0;
                pc[1][k] = pn1;
            }
            __syncwarp();   // the p tile is complete
        }

        // ---- phase 1: gradient, metric, t-part of the divergence; r- and s-fluxes to shared memory (dead tiles) at the
        // lane's own nodes -- the s-fluxes come back as the lane's own A fragment, so only the r-fluxes change hands
        double wc[2][NX];
        double *sws = su;
#pragma unroll
        for (int m = 0; m < NX; m++) wc[0][m] = wc[1][m] = 0.0;
#pragma unroll
        for (int k = 0; k < NX; k++) {
            const double2 b = *reinterpret_cast<const double2 *>(sf + k * N2 + frag);
            double ur0 = 0.0, ur1 = 0.0, us0 = 0.0, us1 = 0.0, ut0 = 0.0, ut1 = 0.0;
            dmma884(ur0, ur1, dA[0], b.x);
            dmma884(ur0, ur1, dA[1], b.y);
            dmma884(us0, us1, pc[0][k], dB[0]);
            dmma884(us0, us1, pc[1][k], dB[1]);
#pragma unroll
            for (int m = 0; m < NX; m++) {
                ut0 = fma(c_D[k * NX + m], pc[0][m], ut0);
                ut1 = fma(c_D[k * NX + m], pc[1][m], ut1);
            }
            // G_ab(i,j,k) = c_ab * w_i w_j w_k
            const double W0 = wij[0] * c_w[k], W1 = wij[1] * c_w[k];
            const double wr0 = fma(c0, ur0, fma(c1, us0, c2 * ut0)) * W0, wr1 = fma(c0, ur1, fma(c1, us1, c2 * ut1)) * W1;
            const double ws0 = fma(c1, ur0, fma(c3, us0, c4 * ut0)) * W0, ws1 = fma(c1, ur1, fma(c3, us1, c4 * ut1)) * W1;
            const double wt0 = fma(c2, ur0, fma(c4, us0, c5 * ut0)) * W0, wt1 = fma(c2, ur1, fma(c4, us1, c5 * ut1)) * W1;
            swr[k * N2 + own0] = wr0;
            swr[k * N2 + own1] = wr1;
            sws[k * N2 + own0] = ws0;
            sws[k * N2 + own1] = ws1;
#pragma unroll
            for (int m = 0; m < NX; m++) {
                wc[0][m] = fma(c_D[k * NX + m], wt0, wc[0][m]);
                wc[1][m] = fma(c_D[k * NX + m], wt1, wc[1][m]);
            }
        }
        __syncwarp();   // the r-flux tile is complete

        // ---- phase 2: D^T wr + D^T ws on top of the register sum, store, p.w
        double *__restrict__ we = w + (size_t)e * N3;
#pragma unroll
        for (int k = 0; k < NX; k++) {
            const double2 b = *reinterpret_cast<const double2 *>(swr + k * N2 + frag);
            double a0 = wc[0][k], a1 = wc[1][k];
            dmma884(a0, a1, dAt[0], b.x);
            dmma884(a0, a1, dAt[1], b.y);
            dmma884(a0, a1, sws[k * N2 + own0], dBt[0]);
            dmma884(a0, a1, sws[k * N2 + own1], dBt[1]);
            we[k * N2 + own0] = a0;
            we[k * N2 + own1] = a1;
            pap = fma(sf[k * N2 + own0], a0, pap);
            pap = fma(sf[k * N2 + own1], a1, pap);
        }
        __syncwarp();   // every lane is done with this stage
        if (lane == 0) {
            const int en = e + STAGES * stride;
            if (en < nel) {
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                issue(stage, en);
            }
        }
    }
    if (pap_out != nullptr) {
        double b = block_reduce(pap, s_red);
        grid_reduce(b, partials, counter, s_red, [=](double tt) { *pap_out = tt; });
    }
}

template <int WARPS, int STAGES>
inline void launch_ax_cg_affine_mma(const double *r, double *p, double *u, double *w, int nel, bool first, double *pap_out)
{
    Ctx &c = ctx();
    using L = AxCgAffMmaSmem<WARPS, STAGES>;
    static_assert(L::bytes <= 227 * 1024, "shared memory of one SM");
    static bool configured = false;
    if (!configured) {
        NEKB_CUDA(cudaFuncSetAttribute(ax_cg_affine_mma_kernel<WARPS, STAGES, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L::bytes));
        NEKB_CUDA(cudaFuncSetAttribute(ax_cg_affine_mma_kernel<WARPS, STAGES, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L::bytes));
        configured = true;
    }
    const int grid = grid_for((nel + WARPS - 1) / WARPS, 1);
    if (first)
        ax_cg_affine_mma_kernel<WARPS, STAGES, true><<<grid, 32 * WARPS, L::bytes, c.stream>>>(
            r, p, u, c.gc.p, w, nel, c.sc.p, c.partials.p, &c.sc.p->counter[0], pap_out);
    else
        ax_cg_affine_mma_kernel<WARPS, STAGES, false><<<grid, 32 * WARPS, L::bytes, c.stream>>>(
            r, p, u, c.gc.p, w, nel, c.sc.p, c.partials.p, &c.sc.p->counter[0], pap_out);
    NEKB_LAUNCHED();
}

// ---------------------------------------------------------------------------------------------- kernel v6 (general geometry, warp per element, DMMA)
// Kernel v5's mapping with the six per-node factors streamed as in ax_cg_kernel (a stage is 9 tiles = 36 KB): each lane
// reads the factors of its own nodes from the stage (12 conflict-free LDS.64 per plane).  Six warps with one stage each fill the
// shared memory of an SM; the kernel is DRAM-bound (2130 cycles of HBM time per element and SM against ~1200 of FP64 +
// shared-memory pipe time).  Measured at E = 262,144 (profiles/r2s_*): 6 x 1 stages 2.19-2.25 ms, 3 x 2 2.57 ms, 2 x 3 3.5 ms,
// ax_cg_kernel<8,3,2> 2.25-2.34 ms.
template <int WARPS, int STAGES>
struct AxCgMmaSmem {
    static constexpr int N3 = 512;
    static constexpr size_t stage_doubles = 9 * (size_t)N3;                    // r, p, u tiles + the six factor tiles
    static constexpr size_t warp_doubles = STAGES * stage_doubles;
    static constexpr size_t bytes = WARPS * warp_doubles * sizeof(double) + WARPS * STAGES * sizeof(uint64_t) + 64 * sizeof(double);
};

template <int WARPS, int STAGES, bool FIRST>
__global__ void __launch_bounds__(32 * WARPS, 1)
    ax_cg_mma_kernel(const double *__restrict__ r, double *__restrict__ p, double *__restrict__ u,
                            const double *__restrict__ gf, double *__restrict__ w, int nel,
                            const CgScalars *__restrict__ sc, double *__restrict__ partials, unsigned *counter, double *pap_out)
{
    using L = AxCgMmaSmem<WARPS, STAGES>;
    constexpr int NX = 8, N2 = 64, N3 = 512;
    constexpr uint32_t T_BYTES = N3 * sizeof(double), G_BYTES = 6 * N3 * sizeof(double);
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double *smem = reinterpret_cast<double *>(smem_raw);
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + WARPS * L::warp_doubles);
    double *s_red = reinterpret_cast<double *>(bars + WARPS * STAGES);

    const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    const int pg = (g >> 1) + 4 * (g & 1);   // pi(g): the j a B-fragment column / C-fragment column g stands for
    double *wbase = smem + wid * L::warp_doubles;
    uint64_t *full = bars + wid * STAGES;

    if (threadIdx.x == 0) {
        for (int q = 0; q < WARPS * STAGES; q++) mbar_init(&bars[q], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    const double alpha = FIRST ? 0.0 : sc->alpha;
    const double beta = FIRST ? 0.0 : sc->work[1] / sc->rtz1;
    // operand fragments of D (step s of the two k = 4 steps of an 8-long contraction)
    double dA[2], dB[2], dAt[2], dBt[2];
#pragma unroll
    for (int s = 0; s < 2; s++) {
        dA[s] = c_D[g * NX + 2 * t + s];           // ur : A(i = g, m = 2t + s) = D(i, m)
        dB[s] = c_D[pg * NX + t + 4 * s];          // us : B(m = t + 4s, j = pi(g)) = D(j, m)
        dAt[s] = c_D[(2 * t + s) * NX + g];        // D^T wr : A(i = g, m = 2t + s) = D(m, i)
        dBt[s] = c_D[(t + 4 * s) * NX + pg];       // D^T ws : B(m = t + 4s, j = pi(g)) = D(m, j)
    }
    const int own0 = t * NX + g, own1 = (t + 4) * NX + g;   // the lane's two nodes in a plane: 32 consecutive words per warp
    const int frag = pg * NX + 2 * t;                      // B fragment (two adjacent words) in a plane

    const int first = blockIdx.x * WARPS + wid, stride = gridDim.x * WARPS;
    auto issue = [&](int stage, int e) {
        double *st = wbase + stage * L::stage_doubles;
        mbar_expect_tx(&full[stage], G_BYTES + (FIRST ? 1 : 3) * T_BYTES);
        bulk_g2s(st + 3 * N3, gf + (size_t)e * 6 * N3, G_BYTES, &full[stage]);
        bulk_g2s(st, r + (size_t)e * N3, T_BYTES, &full[stage]);
        if (!FIRST) {
            bulk_g2s(st + N3, p + (size_t)e * N3, T_BYTES, &full[stage]);
            bulk_g2s(st + 2 * N3, u + (size_t)e * N3, T_BYTES, &full[stage]);
        }
    };
    if (lane == 0) {
#pragma unroll
        for (int s = 0; s < STAGES; s++) {
            const int e = first + s * stride;
            if (e < nel) issue(s, e);
        }
    }

    double pap = 0.0;
    int it = 0;
    for (int e = first; e < nel; e += stride, it++) {
        const int stage = it % STAGES;
        const uint32_t parity = (uint32_t)(it / STAGES) & 1u;
        mbar_wait(&full[stage], parity);
        double *sr = wbase + stage * L::stage_doubles, *sp = sr + N3, *su = sr + 2 * N3;
        const double *sg = sr + 3 * N3;    // g[c][k][j][i], c = rr, rs, rt, ss, st, tt
        const double *sf = FIRST ? sr : sp;    // the tile holding the search direction p of this iteration
        double *swr = FIRST ? sp : sr;         // dead tile that takes the r-fluxes

        // ---- phase 0: u += alpha p_old ; p = r + beta p_old at the lane's own nodes
        double pc[2][NX];
        double *__restrict__ pe = p + (size_t)e * N3;
        if (FIRST) {
#pragma unroll
            for (int k = 0; k < NX; k++) {
                pc[0][k] = sr[k * N2 + own0];
                pc[1][k] = sr[k * N2 + own1];
                pe[k * N2 + own0] = pc[0][k];
                pe[k * N2 + own1] = pc[1][k];
            }
        } else {
            double *__restrict__ ue = u + (size_t)e * N3;
#pragma unroll
            for (int k = 0; k < NX; k++) {
                const double po0 = sp[k * N2 + own0], po1 = sp[k * N2 + own1];
                const double pn0 = fma(beta, po0, sr[k * N2 + own0]), pn1 = fma(beta, po1, sr[k * N2 + own1]);
                ue[k * N2 + own0] = fma(alpha, po0, su[k * N2 + own0]);
                ue[k * N2 + own1] = fma(alpha, po1, su[k * N2 + own1]);
                sp[k * N2 + own0] = pn0;
                sp[k * N2 + own1] = pn1;
                pe[k * N2 + own0] = pn0;
                pe[k * N2 + own1] = pn1;
                pc[0][k] = pn0;
                pc[1][k] = pn1;
            }
            __syncwarp();   // the p tile is complete
        }

        // ---- phase 1: gradient, metric, t-part of the divergence; r- and s-fluxes to shared memory (dead tiles) at the
        // lane's own nodes -- the s-fluxes come back as the lane's own A fragment, so only the r-fluxes change hands
        double wc[2][NX];
        double *sws = su;
#pragma unroll
        for (int m = 0; m < NX; m++) wc[0][m] = wc[1][m] = 0.0;
#pragma unroll
        for (int k = 0; k < NX; k++) {
            const double2 b = *reinterpret_cast<const double2 *>(sf + k * N2 + frag);
            double ur0 = 0.0, ur1 = 0.0, us0 = 0.0, us1 = 0.0, ut0 = 0.0, ut1 = 0.0;
            dmma884(ur0, ur1, dA[0], b.x);
            dmma884(ur0, ur1, dA[1], b.y);
            dmma884(us0, us1, pc[0][k], dB[0]);
            dmma884(us0, us1, pc[1][k], dB[1]);
#pragma unroll
            for (int m = 0; m < NX; m++) {
                ut0 = fma(c_D[k * NX + m], pc[0][m], ut0);
                ut1 = fma(c_D[k * NX + m], pc[1][m], ut1);
            }
            // the six factors of the lane's two nodes of this plane (same products, same order as ax_cg_kernel)
            const int q0 = k * N2 + own0, q1 = k * N2 + own1;
            const double A0 = sg[q0], A1 = sg[N3 + q0], A2 = sg[2 * N3 + q0], A3 = sg[3 * N3 + q0], A4 = sg[4 * N3 + q0], A5 = sg[5 * N3 + q0];
            const double B0 = sg[q1], B1 = sg[N3 + q1], B2 = sg[2 * N3 + q1], B3 = sg[3 * N3 + q1], B4 = sg[4 * N3 + q1], B5 = sg[5 * N3 + q1];
            const double wr0 = fma(A0, ur0, fma(A1, us0, A2 * ut0)), wr1 = fma(B0, ur1, fma(B1, us1, B2 * ut1));
            const double ws0 = fma(A1, ur0, fma(A3, us0, A4 * ut0)), ws1 = fma(B1, ur1, fma(B3, us1, B4 * ut1));
            const double wt0 = fma(A2, ur0, fma(A4, us0, A5 * ut0)), wt1 = fma(B2, ur1, fma(B4, us1, B5 * ut1));
            swr[k * N2 + own0] = wr0;
            swr[k * N2 + own1] = wr1;
            sws[k * N2 + own0] = ws0;
            sws[k * N2 + own1] = ws1;
#pragma unroll
            for (int m = 0; m < NX; m++) {
                wc[0][m] = fma(c_D[k * NX + m], wt0, wc[0][m]);
                wc[1][m] = fma(c_D[k * NX + m], wt1, wc[1][m]);
            }
        }
        __syncwarp();   // the r-flux tile is complete

        // ---- phase 2: D^T wr + D^T ws on top of the register sum, store, p.w
        double *__restrict__ we = w + (size_t)e * N3;
#pragma unroll
        for (int k = 0; k < NX; k++) {
            const double2 b = *reinterpret_cast<const double2 *>(swr + k * N2 + frag);
            double a0 = wc[0][k], a1 = wc[1][k];
            dmma884(a0, a1, dAt[0], b.x);
            dmma884(a0, a1, dAt[1], b.y);
            dmma884(a0, a1, sws[k * N2 + own0], dBt[0]);
            dmma884(a0, a1, sws[k * N2 + own1], dBt[1]);
            we[k * N2 + own0] = a0;
            we[k * N2 + own1] = a1;
            pap = fma(sf[k * N2 + own0], a0, pap);
            pap = fma(sf[k * N2 + own1], a1, pap);
        }
        __syncwarp();   // every lane is done with this stage
        if (lane == 0) {
            const int en = e + STAGES * stride;
            if (en < nel) {
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                issue(stage, en);
            }
        }
    }
    if (pap_out != nullptr) {
        double b = block_reduce(pap, s_red);
        grid_reduce(b, partials, counter, s_red, [=](double tt) { *pap_out = tt; });
    }
}

template <int WARPS, int STAGES>
inline void launch_ax_cg_mma(const double *r, double *p, double *u, double *w, int nel, bool first, double *pap_out)
{
    Ctx &c = ctx();
    using L = AxCgMmaSmem<WARPS, STAGES>;
    static_assert(L::bytes <= 227 * 1024, "shared memory of one SM");
    static bool configured = false;
    if (!configured) {
        NEKB_CUDA(cudaFuncSetAttribute(ax_cg_mma_kernel<WARPS, STAGES, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L::bytes));
        NEKB_CUDA(cudaFuncSetAttribute(ax_cg_mma_kernel<WARPS, STAGES, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L::bytes));
        configured = true;
    }
    const int grid = grid_for((nel + WARPS - 1) / WARPS, 1);
    if (first)
        ax_cg_mma_kernel<WARPS, STAGES, true><<<grid, 32 * WARPS, L::bytes, c.stream>>>(
            r, p, u, c.g.p, w, nel, c.sc.p, c.partials.p, &c.sc.p->counter[0], pap_out);
    else
        ax_cg_mma_kernel<WARPS, STAGES, false><<<grid, 32 * WARPS, L::bytes, c.stream>>>(
            r, p, u, c.g.p, w, nel, c.sc.p, c.partials.p, &c.sc.p->counter[0], pap_out);
    NEKB_LAUNCHED();
}

template <int NX, int GROUPS, int STAGES>
inline void launch_ax_cg(const double *r, double *p, double *u, double *w, int nel, bool first, double *pap_out)
{
    Ctx &c = ctx();
    using L = AxCgSmem<NX, GROUPS, STAGES>;
    static bool configured = false;
    if (!configured) {
        NEKB_CUDA(cudaFuncSetAttribute(ax_cg_kernel<NX, GROUPS, STAGES, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L::bytes));
        NEKB_CUDA(cudaFuncSetAttribute(ax_cg_kernel<NX, GROUPS, STAGES, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L::bytes));
        configured = true;
    }
    const int grid = grid_for((nel + GROUPS - 1) / GROUPS, 1);
    if (first)
        ax_cg_kernel<NX, GROUPS, STAGES, true><<<grid, NX * NX * GROUPS, L::bytes, c.stream>>>(
            r, p, u, c.g.p, w, nel, c.sc.p, c.partials.p, &c.sc.p->counter[0], pap_out);
    else
        ax_cg_kernel<NX, GROUPS, STAGES, false><<<grid, NX * NX * GROUPS, L::bytes, c.stream>>>(
            r, p, u, c.g.p, w, nel, c.sc.p, c.partials.p, &c.sc.p->counter[0], pap_out);
    NEKB_LAUNCHED();
}

// Jacobi diagonal, core/hmholtz.f:380-524 setprec before its dssum + invcol1 (:520-521).
// One CTA per element, one thread per node.
template <int NX>
__global__ void __launch_bounds__(NX *NX *NX)
    setprec_kernel(double *__restrict__ dpc, const double *__restrict__ g, const double *__restrict__ h1,
                   const double *__restrict__ h2, const double *__restrict__ bm1, int nel)
{
    constexpr int N2 = NX * NX, N3 = NX * NX * NX, L = NX - 1;
    __shared__ double s_g[3][N3];
    const int e = blockIdx.x, q = threadIdx.x;
    const int i = q % NX, j = (q / NX) % NX, k = q / N2;
    const double *ge = g + (size_t)e * 6 * N3;
    s_g[0][q] = ge[0 * N3 + q];  // rr
    s_g[1][q] = ge[3 * N3 + q];  // ss
    s_g[2][q] = ge[5 * N3 + q];  // tt
    __syncthreads();
    double a = 0.0;
#pragma unroll
    for (int m = 0; m < NX; m++) a = fma(s_g[0][k * N2 + j * NX + m], c_D[m * NX + i] * c_D[m * NX + i], a);
#pragma unroll
    for (int m = 0; m < NX; m++) a = fma(s_g[1][k * N2 + m * NX + i], c_D[m * NX + j] * c_D[m * NX + j], a);
#pragma unroll
    for (int m = 0; m < NX; m++) a = fma(s_g[2][m * N2 + j * NX + i], c_D[m * NX + k] * c_D[m * NX + k], a);
    // cross terms, added by the reference only at the 8 corners (:440-468); zero factors were stored
    // for undeformed elements, so no flag is needed here.
    if ((i == 0 || i == L) && (j == 0 || j == L) && (k == 0 || k == L)) {
        const double grs = ge[1 * N3 + q], grt = ge[2 * N3 + q], gst = ge[4 * N3 + q];
        const double di = c_D[i * NX + i], dj = c_D[j * NX + j], dk = c_D[k * NX + k];
        a = a + grs * di * dj + grt * di * dk;
        a = a + grs * dj * di + gst * dj * dk;
        a = a + grt * dk * di + gst * dk * dj;
    }
    const size_t o = (size_t)e * N3 + q;
    a = a * h1[o];
    a = fma(h2[o], bm1[o], a);
    dpc[o] = a;
}

template <int NX, int EPB>
inline void launch_ax_t(const double *u, double *w, const double *h1, const double *h2, int nel, double *pap_out)
{
    Ctx &c = ctx();
    const int ngroups = (nel + EPB - 1) / EPB;
    const int grid = grid_for(ngroups, 4);
    c.partials.ensure(4 * 148 * 16);
    unsigned *counter = &c.sc.p->counter[0];
    if (h1 != nullptr)
        ax_kernel<NX, EPB, true><<<grid, NX * NX * EPB, 0, c.stream>>>(u, c.g.p, w, h1, h2, c.bm1.p, nel,
                                                                        c.partials.p, counter, pap_out);
    else
        ax_kernel<NX, EPB, false><<<grid, NX * NX * EPB, 0, c.stream>>>(u, c.g.p, w, nullptr, nullptr, nullptr, nel,
                                                                         c.partials.p, counter, pap_out);
    NEKB_LAUNCHED();
}

// w = A u (h1 == nullptr: pure stiffness as in BP5) ; optional pap_out (device) = sum u.w
template <int NX, int GROUPS, int STAGES>
inline void launch_ax_tma(const double *u, double *w, int nel, double *pap_out)
{
    Ctx &c = ctx();
    using L = AxTmaSmem<NX, GROUPS, STAGES>;
    static bool configured = false;
    if (!configured) {
        NEKB_CUDA(cudaFuncSetAttribute(ax_tma_kernel<NX, GROUPS, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L::bytes));
        configured = true;
    }
    const int ngroups = (nel + GROUPS - 1) / GROUPS;
    const int grid = grid_for(ngroups, 1);
    ax_tma_kernel<NX, GROUPS, STAGES><<<grid, NX * NX * GROUPS, L::bytes, c.stream>>>(u, c.g.p, w, nel, c.partials.p,
                                                                                        &c.sc.p->counter[0], pap_out);
    NEKB_LAUNCHED();
}

inline int ax_variant()
{
    static int v = -1;
    if (v < 0) {
        const char *e = getenv("NEKB_AX_VARIANT");
        v = e ? atoi(e) : 1;
    }
    return v;
}

inline void launch_ax(const double *u, double *w, const double *h1, const double *h2, int nel, double *pap_out)
{
    Ctx &c = ctx();
    NEKB_REQUIRE(c.have_geom && c.have_D, "geometry / derivative matrix not registered");
    if (nel <= 0) {
        if (pap_out) NEKB_CUDA(cudaMemsetAsync(pap_out, 0, sizeof(double), c.stream));
        return;
    }
    if (h1 == nullptr && c.nx == 8 && ax_variant() > 0) {
        switch (ax_variant()) {
            case 2: launch_ax_tma<8, 2, 3>(u, w, nel, pap_out); break;
            case 3: launch_ax_tma<8, 2, 2>(u, w, nel, pap_out); break;
            default: launch_ax_tma<8, 3, 2>(u, w, nel, pap_out); break;
        }
        return;
    }
    switch (c.nx) {
        case 2: launch_ax_t<2, 32>(u, w, h1, h2, nel, pap_out); break;
        case 3: launch_ax_t<3, 8>(u, w, h1, h2, nel, pap_out); break;
        case 4: launch_ax_t<4, 8>(u, w, h1, h2, nel, pap_out); break;
        case 5: launch_ax_t<5, 4>(u, w, h1, h2, nel, pap_out); break;
        case 6: launch_ax_t<6, 4>(u, w, h1, h2, nel, pap_out); break;
        case 7: launch_ax_t<7, 2>(u, w, h1, h2, nel, pap_out); break;
        case 8: launch_ax_t<8, 2>(u, w, h1, h2, nel, pap_out); break;
        case 10: launch_ax_t<10, 1>(u, w, h1, h2, nel, pap_out); break;
        case 12: launch_ax_t<12, 1>(u, w, h1, h2, nel, pap_out); break;
        default: NEKB_REQUIRE(false, "unsupported lx1 (supported: 2-8, 10, 12)");
    }
}

inline void launch_setprec(double *dpc, const double *h1, const double *h2, int nel)
{
    Ctx &c = ctx();
    NEKB_REQUIRE(c.have_geom && c.have_D, "geometry / derivative matrix not registered");
    if (nel <= 0) return;
    switch (c.nx) {
#define NEKB_CASE(NXV)                                                                                         \
    case NXV:                                                                                                  \
        setprec_kernel<NXV><<<nel, NXV * NXV * NXV, 0, c.stream>>>(dpc, c.g.p, h1, h2, c.bm1.p, nel);        \
        break;
        NEKB_CASE(2) NEKB_CASE(3) NEKB_CASE(4) NEKB_CASE(5) NEKB_CASE(6) NEKB_CASE(7) NEKB_CASE(8) NEKB_CASE(10)
#undef NEKB_CASE
        default: NEKB_REQUIRE(false, "unsupported lx1 for setprec (supported: 2-8, 10)");
    }
    NEKB_LAUNCHED();
}

}  // namespace nekb
