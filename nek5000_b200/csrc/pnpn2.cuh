// pnpn2.cuh -- the pressure operator of the Pn-Pn-2 formulation, E = D (h2 B)^-1 D^T:
//   opgradt -> cdtp   core/navier1.f:4095-4114, :330-536   D^T : pressure (Gauss, lx2 = lx1-2) -> velocity (GLL) mesh
//   opdiv   -> multd  core/navier1.f:4064-4093, :538-714   D   : velocity -> pressure mesh
//   opbinv            core/navier1.f:775-850               (h2 B)^-1 with mask + dssum
//   cdabdtp           core/navier1.f:258-293               intype = 1: D (B/dt)^-1 D^T ; 0 / -1: D (h1 A + h2 B)^-1 D^T via ophinv
// 3-D, non-axisymmetric, ifsplit = .false. branch.  One CTA per element: the (lx2)^3 <-> (lx1)^3 tensor contractions with the
// 6x8 interpolation / derivative matrices ixm12, dxm12 run through shared memory; the nine mesh-2 metric arrays (rxm2 ...
// tzm2) are streamed once per element (opgradt: 10 x 216 words in, 3 x 512 out; opdiv: 3 x 512 + 10 x 216 in, 216 out).
#pragma once
#include "proj.cuh"

namespace nekb {

struct Mesh2 {
    bool ready = false;
    int lx2 = 0;
    DevBuf<double> i12, d12;     // [lx2][lx1] row-major: i12[a*lx1+i] = ixm12(a,i)
    DevBuf<double> w3;           // [lx2^3]
    DevBuf<double> met[9];       // rxm2, sxm2, txm2, rym2, sym2, tym2, rzm2, szm2, tzm2 : [isd][q]
    DevBuf<double> bm2, bm2inv, ml, mu;
    double volvm2 = 0.0;
    double tolhs = 1e-8;         // TSTEP tolhs, INPUT nmxv: the velocity solves inside cdabdtp(intype = 0, -1)
    int nmxv = 1000;
    int64_t nelgv = 0;
    bool ifvcor = false;
    double tolps = 1e-8, param21 = 0.0, prelax = 0.0, tolpdf = 0.0;   // TSTEP tolps, INPUT param(21), TSTEP prelax, tolpdf
    // uzawa_gmres storage (core/GMRES: v_gmres, z_gmres, ...)
    std::vector<DevBuf<double>> V, Z;
    DevBuf<double> r, w, x, ones, scal, t1;
    double div0 = 0.0, divex = 0.0;
};
inline Mesh2 &mesh2()
{
    static Mesh2 m;
    return m;
}
struct Met9 {
    const double *p[9];
};

// One term of cdtp: out(i,j,k) = sum_{a,b,c} Ax(a,i) Ay(b,j) Az(c,k) f(a,b,c), f on the N2^3 grid; mxm order x, y, z
template <int N1, int N2>
__device__ __forceinline__ void up_term(const double *f, double *t1, double *t2, const double *Ax, const double *Ay, const double *Az,
                                        double *term, int nthreads)
{
    for (int o = threadIdx.x; o < N1 * N2 * N2; o += nthreads) {  // t1[c][b][i]
        const int i = o % N1, cb = o / N1;
        double s = 0.0;
#pragma unroll
        for (int a = 0; a < N2; a++) s = fma(Ax[a * N1 + i], f[cb * N2 + a], s);
        t1[o] = s;
    }
    __syncthreads();
    for (int o = threadIdx.x; o < N1 * N1 * N2; o += nthreads) {  // t2[c][j][i]
        const int i = o % N1, j = (o / N1) % N1, c = o / (N1 * N1);
        double s = 0.0;
#pragma unroll
        for (int b = 0; b < N2; b++) s = fma(t1[(c * N2 + b) * N1 + i], Ay[b * N1 + j], s);
        t2[o] = s;
    }
    __syncthreads();
    int slot = 0;
    for (int o = threadIdx.x; o < N1 * N1 * N1; o += nthreads, slot++) {
        const int ij = o % (N1 * N1), k = o / (N1 * N1);
        double s = 0.0;
#pragma unroll
        for (int c = 0; c < N2; c++) s = fma(t2[c * N1 * N1 + ij], Az[c * N1 + k], s);
        term[slot] = s;
    }
    __syncthreads();
}

// opgradt: out_isd = sum_q T_q( w3m2 * p * q_isd,m2 ),  T_r = I^T (x) I^T (x) D^T etc.
template <int N1, int N2>
__global__ void __launch_bounds__(256)
    opgradt_kernel(double *__restrict__ ox, double *__restrict__ oy, double *__restrict__ oz, const double *__restrict__ p, Met9 M,
                   const double *__restrict__ w3, const double *__restrict__ i12, const double *__restrict__ d12, int nel)
{
    constexpr int P2 = N2 * N2 * N2, P1 = N1 * N1 * N1, SL = (P1 + 255) / 256;
    __shared__ double sI[N2 * N1], sD[N2 * N1], wx[P2], f[P2], t1[N1 * N2 * N2], t2[N1 * N1 * N2];
    for (int t = threadIdx.x; t < N2 * N1; t += 256) sI[t] = i12[t], sD[t] = d12[t];
    for (int e = blockIdx.x; e < nel; e += gridDim.x) {
        __syncthreads();
        for (int t = threadIdx.x; t < P2; t += 256) wx[t] = w3[t] * p[(size_t)e * P2 + t];   // col3(wx,w3m2,x)
        __syncthreads();
        double *outs[3] = {ox, oy, oz};
#pragma unroll 1
        for (int isd = 0; isd < 3; isd++) {
            double acc[SL], term[SL];
#pragma unroll 1
            for (int q = 0; q < 3; q++) {
                const double *mq = M.p[isd * 3 + q] + (size_t)e * P2;
                for (int t = threadIdx.x; t < P2; t += 256) f[t] = wx[t] * mq[t];
                __syncthreads();
                up_term<N1, N2>(f, t1, t2, q == 0 ? sD : sI, q == 1 ? sD : sI, q == 2 ? sD : sI, term, 256);
#pragma unroll
                for (int sl = 0; sl < SL; sl++) acc[sl] = q == 0 ? term[sl] : acc[sl] + term[sl];   // add2
            }
            int sl = 0;
            for (int o = threadIdx.x; o < P1; o += 256, sl++) outs[isd][(size_t)e * P1 + o] = acc[sl];
        }
    }
}

// One term of multd: v(a,b,c) = sum_{i,j,k} Ax(a,i) Ay(b,j) Az(c,k) u(i,j,k)
template <int N1, int N2>
__device__ __forceinline__ double down_term(const double *u, double *t1, double *t2, const double *Ax, const double *Ay, const double *Az,
                                            int nthreads)
{
    for (int o = threadIdx.x; o < N2 * N1 * N1; o += nthreads) {  // t1[k][j][a]
        const int a = o % N2, jk = o / N2;
        double s = 0.0;
#pragma unroll
        for (int i = 0; i < N1; i++) s = fma(Ax[a * N1 + i], u[jk * N1 + i], s);
        t1[o] = s;
    }
    __syncthreads();
    for (int o = threadIdx.x; o < N2 * N2 * N1; o += nthreads) {  // t2[k][b][a]
        const int a = o % N2, b = (o / N2) % N2, k = o / (N2 * N2);
        double s = 0.0;
#pragma unroll
        for (int j = 0; j < N1; j++) s = fma(t1[(k * N1 + j) * N2 + a], Ay[b * N1 + j], s);
        t2[o] = s;
    }
    __syncthreads();
    double v = 0.0;
    if (threadIdx.x < N2 * N2 * N2) {
        const int ab = threadIdx.x % (N2 * N2), c = threadIdx.x / (N2 * N2);
#pragma unroll
        for (int k = 0; k < N1; k++) v = fma(t2[k * N2 * N2 + ab], Az[c * N1 + k], v);
    }
    __syncthreads();
    return v;
}

// opdiv: out = sum_isd w3m2 * sum_q q_isd,m2 * T_q^T( u_isd )
template <int N1, int N2>
__global__ void __launch_bounds__(256)
    opdiv_kernel(double *__restrict__ out, const double *__restrict__ ux, const double *__restrict__ uy, const double *__restrict__ uz, Met9 M,
                 const double *__restrict__ w3, const double *__restrict__ i12, const double *__restrict__ d12, int nel)
{
    constexpr int P2 = N2 * N2 * N2, P1 = N1 * N1 * N1;
    static_assert(P2 <= 256, "one pressure node per thread");
    __shared__ double sI[N2 * N1], sD[N2 * N1], u[P1], t1[N2 * N1 * N1], t2[N2 * N2 * N1];
    for (int t = threadIdx.x; t < N2 * N1; t += 256) sI[t] = i12[t], sD[t] = d12[t];
    for (int e = blockIdx.x; e < nel; e += gridDim.x) {
        const double *us[3] = {ux, uy, uz};
        double tot = 0.0;
#pragma unroll 1
        for (int isd = 0; isd < 3; isd++) {
            __syncthreads();
            for (int t = threadIdx.x; t < P1; t += 256) u[t] = us[isd][(size_t)e * P1 + t];
            __syncthreads();
            double dx = 0.0;
#pragma unroll 1
            for (int q = 0; q < 3; q++) {
                const double v = down_term<N1, N2>(u, t1, t2, q == 0 ? sD : sI, q == 1 ? sD : sI, q == 2 ? sD : sI, 256);
                if (threadIdx.x < P2) {
                    const double mq = M.p[isd * 3 + q][(size_t)e * P2 + threadIdx.x];
                    dx = q == 0 ? v * mq : dx + v * mq;                                    // col2 / addcol3
                }
            }
            if (threadIdx.x < P2) {
                dx = dx * w3[threadIdx.x];                                                 // col2(dx,w3m2)
                tot = isd == 0 ? dx : tot + dx;                                            // copy / add2
            }
        }
        if (threadIdx.x < P2) out[(size_t)e * P2 + threadIdx.x] = tot;
    }
}

// out_i = inp_i / dssum(bm1 / h2inv)  after  inp_i <- dssum(mask_i inp_i)   (opbinv; inp is modified as in the reference)
__global__ void __launch_bounds__(256)
    uz_split_kernel(double *__restrict__ ml, double *__restrict__ mu, const double *__restrict__ b, const double *__restrict__ binv, int64_t n)
{   // core/gmres.f:240-252 uzawa_gmres_split0
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < n; t += (int64_t)gridDim.x * blockDim.x)
        ml[t] = sqrt(binv[t]), mu[t] = sqrt(b[t]);
}
// a = m * (b - c)  (c == nullptr: a = m * b)
__global__ void __launch_bounds__(256)
    uz_resid_kernel(double *__restrict__ a, const double *__restrict__ m, const double *__restrict__ b, const double *__restrict__ c, int64_t n)
{
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < n; t += (int64_t)gridDim.x * blockDim.x)
        a[t] = m[t] * (c ? b[t] - c[t] : b[t]);
}
__global__ void __launch_bounds__(256) invcol3_kernel(double *__restrict__ o, const double *__restrict__ a, const double *__restrict__ b, int64_t n)
{
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < n; t += (int64_t)gridDim.x * blockDim.x) o[t] = a[t] / b[t];
}
__global__ void __launch_bounds__(256)
    opbinv_fin_kernel(double *__restrict__ o1, double *__restrict__ o2, double *__restrict__ o3, const double *__restrict__ i1,
                      const double *__restrict__ i2, const double *__restrict__ i3, int64_t n)
{
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < n; t += (int64_t)gridDim.x * blockDim.x) {
        const double tmp = 1.0 / o1[t];
        o1[t] = i1[t] * tmp;
        o2[t] = i2[t] * tmp;
        o3[t] = i3[t] * tmp;
    }
}

inline Met9 met9()
{
    Met9 m;
    for (int k = 0; k < 9; k++) m.p[k] = mesh2().met[k].p;
    return m;
}
inline void require_mesh2()
{
    Ctx &c = ctx();
    NEKB_REQUIRE(mesh2().ready, "Pn-Pn-2 operators: mesh-2 data not registered (nekb_set_mesh2)");
    NEKB_REQUIRE(c.nx == 8 && mesh2().lx2 == 6, "Pn-Pn-2 operators are built for lx1 = 8, lx2 = 6");
}
inline void opgradt_dev(double *ox, double *oy, double *oz, const double *p)
{
    Ctx &c = ctx();
    require_mesh2();
    if (c.nelv <= 0) return;
    opgradt_kernel<8, 6><<<grid_for(c.nelv, 4), 256, 0, c.stream>>>(ox, oy, oz, p, met9(), mesh2().w3.p, mesh2().i12.p, mesh2().d12.p, c.nelv);
    NEKB_LAUNCHED();
}
inline void opdiv_dev(double *out, const double *ux, const double *uy, const double *uz)
{
    Ctx &c = ctx();
    require_mesh2();
    if (c.nelv <= 0) return;
    opdiv_kernel<8, 6><<<grid_for(c.nelv, 4), 256, 0, c.stream>>>(out, ux, uy, uz, met9(), mesh2().w3.p, mesh2().i12.p, mesh2().d12.p, c.nelv);
    NEKB_LAUNCHED();
}
inline void opbinv_dev(double *o1, double *o2, double *o3, double *i1, double *i2, double *i3, const double *h2inv, int gs_handle)
{
    Ctx &c = ctx();
    const int64_t n = (int64_t)c.nelv * c.nxyz;
    NEKB_REQUIRE(c.vmask[0].n >= (size_t)n && c.vmask[1].n >= (size_t)n && c.vmask[2].n >= (size_t)n,
                 "opbinv: v1mask, v2mask, v3mask not registered (nekb_set_velocity_state)");
    NEKB_REQUIRE(c.bm1.n >= (size_t)n, "opbinv: bm1 not registered");
    double *in[3] = {i1, i2, i3};
    for (int k = 0; k < 3; k++) {                                 // opmask + opdssum
        col2_kernel<<<cg_grid(n), 256, 0, c.stream>>>(in[k], c.vmask[k].p, n);
        NEKB_LAUNCHED();
        gs_op(gs_handle, in[k], 1, nullptr);
    }
    invcol3_kernel<<<cg_grid(n), 256, 0, c.stream>>>(o1, c.bm1.p, h2inv, n);
    NEKB_LAUNCHED();
    gs_op(gs_handle, o1, 1, nullptr);
    opbinv_fin_kernel<<<cg_grid(n), 256, 0, c.stream>>>(o1, o2, o3, i1, i2, i3, n);
    NEKB_LAUNCHED();
}

struct Pnpn2Work {
    DevBuf<double> ta[3], tb[3];
};
inline Pnpn2Work &pnpn2_work()
{
    static Pnpn2Work w;
    return w;
}
}  // namespace nekb
