// fdm_h1.cuh -- single-level overlapping-Schwarz / fast-diagonalisation preconditioner on the lx1^3 tiles.
//
// Replaces core/hmholtz.f:937-1026 fdm_h1, :1028-1112 set_fdm_prec_h1A_gen (nine generic 1-D eigen-systems: the two
// ends of a direction are each Internal / Dirichlet / Neumann), :1114-1220 set_fdm_prec_h1A_els (ktype, elsize) and
// :1222-1290 set_fdm_prec_h1b (the diagonal).  Used by cggo's Schwarz branch (kfldfdm >= 0, hmholtz.f:731-746) and by
// hmh_gmres when ifmgrid = .false. (gmres.f:412-420).
//
// One CTA pass per element: S^T (x) S^T (x) S^T, the diagonal, S (x) S (x) S, all in shared memory; the nine S
// matrices (9*lx1^2 doubles) are a table that stays cache resident.  Then dssum + mask through gs_op.
#pragma once
#include "hsmg.cuh"

namespace nekb {

struct FdmH1State {
    bool ready = false;
    int nel = 0;
    DevBuf<double> fds;       // [9][nx*nx], S[i*nx+a] (eigenvectors in columns)
    DevBuf<double> dd;        // [9][nx]
    DevBuf<int32_t> ktype;    // [nel][3], 0-based table rows
    DevBuf<double> elsize;    // [nel][3]
    std::vector<int32_t> ktype_host;
    std::vector<double> elsize_host, dd_host, fds_host;
    int kfldfdm = -1;         // < 0: cggo uses the Jacobi preconditioner (core/FDMH1 kfldfdm)
    DevBuf<double> d;         // diagonal of the current solve (cggo)
    DevBuf<double> z;
};
inline FdmH1State &fdm_h1_state()
{
    static FdmH1State s;
    return s;
}

// face_internal[6*nel]: 1 where cbc is 'E  ','P  ','p  ' (symmetric face order r-,r+,s-,s+,t-,t+)
inline void fdm_h1_setup(const int *face_internal, const double *mask, const double *xm1, const double *ym1, const double *zm1,
                         int nel)
{
    Ctx &c = ctx();
    FdmH1State &F = fdm_h1_state();
    ensure_operators();
    const int n = c.nx;
    const int64_t n3 = (int64_t)n * n * n;
    const std::vector<double> &z = c.z_host, &w = c.w_host, &D = c.D_host;  // D[a*n+b] = D(a,b)
    const double delta = fabs(z[1] - z[0]), bbh = 0.5 * delta, aah = 1.0 / delta;
    F.dd_host.assign((size_t)9 * n, 0.0);
    F.fds_host.assign((size_t)9 * n * n, 0.0);
    int l = 0;
    for (int right = 1; right <= 3; right++)
        for (int left = 1; left <= 3; left++, l++) {
            std::vector<double> aa((size_t)n * n, 0.0), bb((size_t)n * n, 0.0), S, lam;
            for (int i = 0; i < n; i++) bb[(size_t)i * n + i] = w[i];
            for (int i = 0; i < n; i++)
                for (int j = 0; j < n; j++) {
                    double s = 0.0;
                    for (int k = 0; k < n; k++) s = s + D[(size_t)k * n + i] * (w[k] * D[(size_t)k * n + j]);
                    aa[(size_t)i * n + j] = s;
                }
            auto fix = [&](int e, int kind) {
                if (kind == 1) {
                    bb[(size_t)e * n + e] += bbh;
                    aa[(size_t)e * n + e] += aah;
                } else if (kind == 2) {
                    bb[(size_t)e * n + e] = 1.0;
                    for (int i = 0; i < n; i++) aa[(size_t)i * n + e] = 0.0, aa[(size_t)e * n + i] = 0.0;
                    aa[(size_t)e * n + e] = 1.0;
                }
            };
            fix(0, left);
            fix(n - 1, right);
            generalev_host(n, aa, bb, S, lam);
            std::copy(S.begin(), S.end(), F.fds_host.begin() + (size_t)l * n * n);
            std::copy(lam.begin(), lam.end(), F.dd_host.begin() + (size_t)l * n);
        }
    F.ktype_host.assign((size_t)3 * nel, 0);
    F.elsize_host.assign((size_t)3 * nel, 0.0);
    auto at = [n](int i, int j, int k) { return (int64_t)i + n * (j + (int64_t)n * k); };
    for (int64_t e = 0; e < nel; e++) {
        const double *me = mask + e * n3, *x = xm1 + e * n3, *y = ym1 + e * n3, *zc = zm1 + e * n3;
        for (int d = 0; d < 3; d++) {
            // hmholtz.f:1166-1180: probe points (1,2,2)/(lx1,2,2), (2,1,2)/(2,lx1,2), (2,2,1)/(2,2,lx1)
            const int64_t k1 = d == 0 ? at(0, 1, 1) : (d == 1 ? at(1, 0, 1) : at(1, 1, 0));
            const int64_t k2 = d == 0 ? at(n - 1, 1, 1) : (d == 1 ? at(1, n - 1, 1) : at(1, 1, n - 1));
            const int ic1 = face_internal[e * 6 + 2 * d] ? 1 : (me[k1] == 0.0 ? 2 : 3);
            const int jc1 = face_internal[e * 6 + 2 * d + 1] ? 1 : (me[k2] == 0.0 ? 2 : 3);
            F.ktype_host[(size_t)e * 3 + d] = ic1 + 3 * (jc1 - 1) - 1;
            // :1222-1252 element size from the two faces normal to d
            double dlm = 0.0, wgt = 0.0;
            for (int b = 0; b < n; b++)
                for (int a = 0; a < n; a++) {
                    const int64_t p0 = d == 0 ? at(0, a, b) : (d == 1 ? at(a, 0, b) : at(a, b, 0));
                    const int64_t p1 = d == 0 ? at(n - 1, a, b) : (d == 1 ? at(a, n - 1, b) : at(a, b, n - 1));
                    const double dx = x[p1] - x[p0], dy = y[p1] - y[p0], dz = zc[p1] - zc[p0];
                    const double dl2 = dx * dx + dy * dy + dz * dz;
                    const double ww = w[0] * w[a] * w[b];
                    dlm = dlm + dl2 * ww;
                    wgt = wgt + ww;
                }
            F.elsize_host[(size_t)e * 3 + d] = sqrt(dlm / wgt) / 2.0;
        }
    }
    F.fds.upload(F.fds_host.data(), F.fds_host.size(), c.stream);
    F.dd.upload(F.dd_host.data(), F.dd_host.size(), c.stream);
    F.ktype.upload(F.ktype_host.data(), F.ktype_host.size(), c.stream);
    F.elsize.upload(F.elsize_host.data(), F.elsize_host.size(), c.stream);
    NEKB_CUDA(cudaStreamSynchronize(c.stream));
    F.nel = nel;
    F.ready = true;
}

inline int fdm_h1_kfldfdm() { return fdm_h1_state().kfldfdm; }

// set_fdm_prec_h1b (hmholtz.f:1222-1290, 3-D, ifbhalf = .false.): one CTA per element
__global__ void __launch_bounds__(256)
    fdm_h1b_kernel(double *__restrict__ d, const double *__restrict__ h1, const double *__restrict__ h2, const double *__restrict__ dd,
                   const int32_t *__restrict__ ktype, const double *__restrict__ elsize, int nx)
{
    __shared__ double red[33];
    __shared__ double s_h[2];
    const int e = blockIdx.x, n3 = nx * nx * nx;
    double a = 0.0, b = 0.0;
    for (int q = threadIdx.x; q < n3; q += blockDim.x) {
        a += h1[(size_t)e * n3 + q];
        b += h2 ? h2[(size_t)e * n3 + q] : 0.0;
    }
    const double ta = block_reduce(a, red);
    const double tb = block_reduce(b, red);
    if (threadIdx.x == 0) s_h[0] = ta / n3, s_h[1] = tb / n3;
    __syncthreads();
    const double h1b = s_h[0], h2b = s_h[1];
    const double s0 = elsize[(size_t)e * 3], s1 = elsize[(size_t)e * 3 + 1], s2 = elsize[(size_t)e * 3 + 2];
    const double vol = s0 * s1 * s2, vl1 = s1 * s2 / s0, vl2 = s0 * s2 / s1, vl3 = s0 * s1 / s2;
    const double *d1 = dd + (size_t)ktype[(size_t)e * 3] * nx, *d2 = dd + (size_t)ktype[(size_t)e * 3 + 1] * nx,
                 *d3 = dd + (size_t)ktype[(size_t)e * 3 + 2] * nx;
    for (int q = threadIdx.x; q < n3; q += blockDim.x) {
        const int i = q % nx, j = (q / nx) % nx, k = q / (nx * nx);
        const double den = h1b * (vl1 * d1[i] + vl2 * d2[j] + vl3 * d3[k]) + h2b * vol;
        d[(size_t)e * n3 + q] = den != 0.0 ? 1.0 / den : 0.0;
    }
}

// fdm_h1 before its dssum (hmholtz.f:971-988): z = (S3 x S2 x S1) d .* (S3^T x S2^T x S1^T) r
template <int NX, int EPB>
__global__ void __launch_bounds__(NX *NX *EPB)
    fdm_h1_kernel(double *__restrict__ z, const double *__restrict__ r, const double *__restrict__ d, const double *__restrict__ fds,
                  const int32_t *__restrict__ ktype, int nel)
{
    constexpr int NXP = (NX % 2 == 0) ? NX + 1 : NX, N2 = NX * NX, N3 = NX * NX * NX, TILE = NX * NX * NXP;
    __shared__ double s_t[EPB][TILE];
    __shared__ double s_S[EPB][3][N2];
    const int es = threadIdx.x / N2, tl = threadIdx.x % N2, p = tl % NX, q = tl / NX;
    const int el = blockIdx.x * EPB + es;
    const bool act = el < nel;
    double *T = s_t[es];
    auto at = [](int i, int j, int k) { return (k * NX + j) * NXP + i; };
    if (act) {
        for (int dir = 0; dir < 3; dir++) s_S[es][dir][tl] = fds[(size_t)ktype[(size_t)el * 3 + dir] * N2 + tl];
        const double *re = r + (size_t)el * N3;
#pragma unroll
        for (int k = 0; k < NX; k++) T[at(p, q, k)] = re[k * N2 + tl];
    }
    __syncthreads();
    double in[NX], out[NX];
#pragma unroll
    for (int dir = 0; dir < 3; dir++) {
        if (act) {
            const int base = dir == 0 ? at(0, p, q) : (dir == 1 ? at(p, 0, q) : at(p, q, 0));
            const int stride = dir == 0 ? 1 : (dir == 1 ? NXP : NX * NXP);
            const double *S = s_S[es][dir];
#pragma unroll
            for (int i = 0; i < NX; i++) in[i] = T[base + i * stride];
#pragma unroll
            for (int a = 0; a < NX; a++) {
                double s = 0.0;
#pragma unroll
                for (int i = 0; i < NX; i++) s = fma(S[i * NX + a], in[i], s);
                out[a] = s;
            }
            if (dir == 2) {
                const double *de = d + (size_t)el * N3;
#pragma unroll
                for (int a = 0; a < NX; a++) out[a] *= de[a * N2 + tl];
            }
#pragma unroll
            for (int a = 0; a < NX; a++) T[base + a * stride] = out[a];
        }
        __syncthreads();
    }
#pragma unroll
    for (int dd_ = 0; dd_ < 3; dd_++) {
        const int dir = 2 - dd_;
        if (act) {
            const int base = dir == 0 ? at(0, p, q) : (dir == 1 ? at(p, 0, q) : at(p, q, 0));
            const int stride = dir == 0 ? 1 : (dir == 1 ? NXP : NX * NXP);
            const double *S = s_S[es][dir];
#pragma unroll
            for (int i = 0; i < NX; i++) in[i] = T[base + i * stride];
#pragma unroll
            for (int a = 0; a < NX; a++) {
                double s = 0.0;
#pragma unroll
                for (int i = 0; i < NX; i++) s = fma(S[a * NX + i], in[i], s);
                out[a] = s;
            }
#pragma unroll
            for (int a = 0; a < NX; a++) T[base + a * stride] = out[a];
        }
        __syncthreads();
    }
    if (act) {
        double *ze = z + (size_t)el * N3;
#pragma unroll
        for (int k = 0; k < NX; k++) ze[k * N2 + tl] = T[at(p, q, k)];
    }
}

inline void set_fdm_prec_h1b_dev(double *d, const double *h1, const double *h2, int nel)
{
    Ctx &c = ctx();
    FdmH1State &F = fdm_h1_state();
    NEKB_REQUIRE(F.ready && nel <= F.nel, "set_fdm_prec_h1b: nekb_fdm_h1_setup has not been called");
    if (nel <= 0) return;
    fdm_h1b_kernel<<<nel, 256, 0, c.stream>>>(d, h1, h2, F.dd.p, F.ktype.p, F.elsize.p, c.nx);
    NEKB_LAUNCHED();
}

template <int NX>
inline void launch_fdm_h1_t(double *z, const double *r, const double *d, int nel)
{
    constexpr int EPB = (NX * NX >= 100) ? 2 : (NX * NX >= 64 ? 4 : 8);
    FdmH1State &F = fdm_h1_state();
    fdm_h1_kernel<NX, EPB><<<(nel + EPB - 1) / EPB, NX * NX * EPB, 0, ctx().stream>>>(z, r, d, F.fds.p, F.ktype.p, nel);
    NEKB_LAUNCHED();
}

// z = mask * dssum( FDM(r) )   (hmholtz.f:937-1026 with ifbhalf = .false.)
inline void fdm_h1_apply(double *z, const double *r, const double *d, const double *mask, int nel, int gs_handle)
{
    Ctx &c = ctx();
    FdmH1State &F = fdm_h1_state();
    NEKB_REQUIRE(F.ready && nel <= F.nel, "fdm_h1: nekb_fdm_h1_setup has not been called");
    if (nel > 0) {
        switch (c.nx) {
            case 4: launch_fdm_h1_t<4>(z, r, d, nel); break;
            case 5: launch_fdm_h1_t<5>(z, r, d, nel); break;
            case 6: launch_fdm_h1_t<6>(z, r, d, nel); break;
            case 7: launch_fdm_h1_t<7>(z, r, d, nel); break;
            case 8: launch_fdm_h1_t<8>(z, r, d, nel); break;
            case 10: launch_fdm_h1_t<10>(z, r, d, nel); break;
            case 12: launch_fdm_h1_t<12>(z, r, d, nel); break;
            default: NEKB_REQUIRE(false, "fdm_h1: unsupported lx1 (supported: 4-8, 10, 12)");
        }
    }
    gs_op(gs_handle, z, 1, mask);
}

}  // namespace nekb
