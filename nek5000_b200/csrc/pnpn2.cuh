// pnpn2.cuh -- the pressure operator of the Pn-Pn-2 formulation, E = D (h2 B)^-1 D^T:
//   opgradt -> cdtp   core/navier1.f:4095-4114, :330-536   D^T : pressure (Gauss, lx2 = lx1-2) -> velocity (GLL) mesh
//   opdiv   -> multd  core/navier1.f:4064-4093, :538-714   D   : velocity -> pressure mesh
//   opbinv            core/navier1.f:775-850               (h2 B)^-1 with mask + dssum
//   cdabdtp           core/navier1.f:258-293               intype = 1: D (B/dt)^-1 D^T ; 0 / -1: D (h1 A + h2 B)^-1 D^T via ophinv
// 3-D, non-axisymmetric, ifsplit = .false. branch.  One WARP per element: the (lx2)^3 <-> (lx1)^3 tensor contractions with the
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

// Warp-per-element kernels.  A warp owns one element at a time (8 warps per CTA); every lane owns whole LINES of a tensor
// contraction: it loads the 6 (or 8) inputs of a line once and produces its 8 (or 6) outputs with the interpolation /
// derivative matrices read from the constant bank as immediate FMA operands (fully unrolled), so a line costs 6 + 8 shared-
// memory accesses for 48 FMAs instead of 96.  Stages of a contraction are separated by __syncwarp only.
constexpr int PN_WARPS = 8;
__constant__ double c_I12[48], c_D12[48];   // [a*8 + i] = ixm12(a,i), dxm12(a,i)  (lx2 = 6, lx1 = 8)

template <bool DER>
__device__ __forceinline__ double m12(int a, int i)
{
    return DER ? c_D12[a * 8 + i] : c_I12[a * 8 + i];
}
// 6 -> 8 along one line: out[i] = sum_a A(a,i) in[a]
template <bool DER>
__device__ __forceinline__ void line_up(const double (&in)[6], double (&out)[8])
{
#pragma unroll
    for (int i = 0; i < 8; i++) {
        double s = 0.0;
#pragma unroll
        for (int a = 0; a < 6; a++) s = fma(m12<DER>(a, i), in[a], s);
        out[i] = s;
    }
}
// 8 -> 6 along one line: out[a] = sum_i A(a,i) in[i]
template <bool DER>
__device__ __forceinline__ void line_down(const double (&in)[8], double (&out)[6])
{
#pragma unroll
    for (int a = 0; a < 6; a++) {
        double s = 0.0;
#pragma unroll
        for (int i = 0; i < 8; i++) s = fma(m12<DER>(a, i), in[i], s);
        out[a] = s;
    }
}

constexpr int FP = 7;   // pitch of the 6-wide rows of f (bank-conflict-free line loads)
constexpr int UP = 9;   // pitch of the 8-wide rows of u

// One term of cdtp (mxm order x, y, z): term(i,j,k) = sum_abc Ax(a,i) Ay(b,j) Az(c,k) f(a,b,c); Q = direction of the derivative.
// f[cb*FP + a]; t1[(c*6+b)*8 + i]; t2[(c*8+j)*8 + i]; term slot 8*l2 + k <-> node (ij = lane + 32 l2, k)
template <int Q>
__device__ __forceinline__ void up_term(const double *f, double *t1, double *t2, double (&term)[16], int lane)
{
    for (int cb = lane; cb < 36; cb += 32) {
        double in[6], out[8];
#pragma unroll
        for (int a = 0; a < 6; a++) in[a] = f[cb * FP + a];
        line_up<Q == 0>(in, out);
#pragma unroll
        for (int i = 0; i < 8; i++) t1[cb * 8 + i] = out[i];
    }
    __syncwarp();
    for (int ci = lane; ci < 48; ci += 32) {
        const int i = ci % 8, c = ci / 8;
        double in[6], out[8];
#pragma unroll
        for (int b = 0; b < 6; b++) in[b] = t1[(c * 6 + b) * 8 + i];
        line_up<Q == 1>(in, out);
#pragma unroll
        for (int j = 0; j < 8; j++) t2[(c * 8 + j) * 8 + i] = out[j];
    }
    __syncwarp();
#pragma unroll
    for (int l2 = 0; l2 < 2; l2++) {
        const int ij = lane + 32 * l2;
        double in[6], out[8];
#pragma unroll
        for (int c = 0; c < 6; c++) in[c] = t2[c * 64 + ij];
        line_up<Q == 2>(in, out);
#pragma unroll
        for (int k = 0; k < 8; k++) term[8 * l2 + k] = out[k];
    }
    __syncwarp();
}

// opgradt: out_isd = sum_q T_q( w3m2 * p * q_isd,m2 ),  T_r = I^T (x) I^T (x) D^T etc.
__global__ void __launch_bounds__(32 * PN_WARPS, 1)
    opgradt_kernel(double *__restrict__ ox, double *__restrict__ oy, double *__restrict__ oz, const double *__restrict__ p, Met9 M,
                   const double *__restrict__ w3, int nel)
{
    constexpr int P2 = 216, P1 = 512, PER = P2 + 36 * FP + 288 + 384;
    extern __shared__ double sm_up[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double *wx = sm_up + warp * PER, *f = wx + P2, *t1 = f + 36 * FP, *t2 = t1 + 288;
    double *outs[3] = {ox, oy, oz};
    for (int e = blockIdx.x * PN_WARPS + warp; e < nel; e += gridDim.x * PN_WARPS) {
        __syncwarp();
        for (int t = lane; t < P2; t += 32) wx[t] = w3[t] * p[(size_t)e * P2 + t];   // col3(wx,w3m2,x)
        // the metric tile of the NEXT (isd, q) term is fetched into registers while the current term is contracted
        double mreg[7];
#pragma unroll
        for (int s7 = 0; s7 < 7; s7++) {
            const int t = lane + 32 * s7;
            mreg[s7] = t < P2 ? M.p[0][(size_t)e * P2 + t] : 0.0;
        }
        __syncwarp();
#pragma unroll 1
        for (int isd = 0; isd < 3; isd++) {
            double acc[16], term[16];
#pragma unroll
            for (int q = 0; q < 3; q++) {
#pragma unroll
                for (int s7 = 0; s7 < 7; s7++) {
                    const int t = lane + 32 * s7;
                    if (t < P2) f[(t / 6) * FP + t % 6] = wx[t] * mreg[s7];
                }
                const int nxt = isd * 3 + q + 1;
                if (nxt < 9) {
                    const double *mq = M.p[nxt] + (size_t)e * P2;
#pragma unroll
                    for (int s7 = 0; s7 < 7; s7++) {
                        const int t = lane + 32 * s7;
                        mreg[s7] = t < P2 ? mq[t] : 0.0;
                    }
                }
                __syncwarp();
                if (q == 0) up_term<0>(f, t1, t2, term, lane);
                if (q == 1) up_term<1>(f, t1, t2, term, lane);
                if (q == 2) up_term<2>(f, t1, t2, term, lane);
#pragma unroll
                for (int sl = 0; sl < 16; sl++) acc[sl] = q == 0 ? term[sl] : acc[sl] + term[sl];   // add2
            }
            double *o = outs[isd] + (size_t)e * P1;
#pragma unroll
            for (int l2 = 0; l2 < 2; l2++)
#pragma unroll
                for (int k = 0; k < 8; k++) o[k * 64 + lane + 32 * l2] = acc[8 * l2 + k];
        }
    }
}

// One term of multd: v(a,b,c) = sum_ijk Ax(a,i) Ay(b,j) Az(c,k) u(i,j,k).  u[jk*UP + i]; t1[(k*8+j)*6 + a]; t2[(k*6+b)*6 + a];
// v slot 6*l2 + c <-> node (ab = lane + 32 l2 < 36, c)
template <int Q>
__device__ __forceinline__ void down_term(const double *u, double *t1, double *t2, double (&v)[12], int lane)
{
#pragma unroll
    for (int l2 = 0; l2 < 2; l2++) {
        const int jk = lane + 32 * l2;
        double in[8], out[6];
#pragma unroll
        for (int i = 0; i < 8; i++) in[i] = u[jk * UP + i];
        line_down<Q == 0>(in, out);
#pragma unroll
        for (int a = 0; a < 6; a++) t1[jk * 6 + a] = out[a];
    }
    __syncwarp();
    for (int ka = lane; ka < 48; ka += 32) {
        const int a = ka % 6, k = ka / 6;
        double in[8], out[6];
#pragma unroll
        for (int j = 0; j < 8; j++) in[j] = t1[(k * 8 + j) * 6 + a];
        line_down<Q == 1>(in, out);
#pragma unroll
        for (int b = 0; b < 6; b++) t2[(k * 6 + b) * 6 + a] = out[b];
    }
    __syncwarp();
#pragma unroll
    for (int l2 = 0; l2 < 2; l2++) {
        const int ab = lane + 32 * l2;
        double in[8], out[6];
        if (ab < 36) {
#pragma unroll
            for (int k = 0; k < 8; k++) in[k] = t2[k * 36 + ab];
            line_down<Q == 2>(in, out);
        } else {
#pragma unroll
            for (int c = 0; c < 6; c++) out[c] = 0.0;
        }
#pragma unroll
        for (int c = 0; c < 6; c++) v[6 * l2 + c] = out[c];
    }
    __syncwarp();
}

// opdiv: out = sum_isd w3m2 * sum_q q_isd,m2 * T_q^T( u_isd )
__global__ void __launch_bounds__(32 * PN_WARPS, 1)
    opdiv_kernel(double *__restrict__ out, const double *__restrict__ ux, const double *__restrict__ uy, const double *__restrict__ uz, Met9 M,
                 const double *__restrict__ w3, int nel)
{
    constexpr int P2 = 216, P1 = 512, PER = 64 * UP + 384 + 288;
    extern __shared__ double sm_dn[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double *u = sm_dn + warp * PER, *t1 = u + 64 * UP, *t2 = t1 + 384;
    const double *us[3] = {ux, uy, uz};
    for (int e = blockIdx.x * PN_WARPS + warp; e < nel; e += gridDim.x * PN_WARPS) {
        double tot[12], ureg[16];
#pragma unroll
        for (int s = 0; s < 16; s++) ureg[s] = us[0][(size_t)e * P1 + lane + 32 * s];
#pragma unroll 1
        for (int isd = 0; isd < 3; isd++) {
            __syncwarp();
#pragma unroll
            for (int s = 0; s < 16; s++) {
                const int t = lane + 32 * s;
                u[(t / 8) * UP + t % 8] = ureg[s];
            }
            if (isd < 2) {  // the next component streams in while this one is contracted
#pragma unroll
                for (int s = 0; s < 16; s++) ureg[s] = us[isd + 1][(size_t)e * P1 + lane + 32 * s];
            }
            double mreg[3][12];  // the three metric tiles of this component
#pragma unroll
            for (int q = 0; q < 3; q++) {
                const double *mq = M.p[isd * 3 + q] + (size_t)e * P2;
#pragma unroll
                for (int l2 = 0; l2 < 2; l2++)
#pragma unroll
                    for (int c = 0; c < 6; c++) {
                        const int ab = lane + 32 * l2;
                        mreg[q][6 * l2 + c] = ab < 36 ? mq[c * 36 + ab] : 0.0;
                    }
            }
            __syncwarp();
            double dx[12], v[12];
#pragma unroll
            for (int q = 0; q < 3; q++) {
                if (q == 0) down_term<0>(u, t1, t2, v, lane);
                if (q == 1) down_term<1>(u, t1, t2, v, lane);
                if (q == 2) down_term<2>(u, t1, t2, v, lane);
#pragma unroll
                for (int sl = 0; sl < 12; sl++) dx[sl] = q == 0 ? v[sl] * mreg[q][sl] : dx[sl] + v[sl] * mreg[q][sl];   // col2 / addcol3
            }
#pragma unroll
            for (int l2 = 0; l2 < 2; l2++)
#pragma unroll
                for (int c = 0; c < 6; c++) {
                    const int ab = lane + 32 * l2;
                    const double d = dx[6 * l2 + c] * (ab < 36 ? w3[c * 36 + ab] : 0.0);                    // col2(dx,w3m2)
                    tot[6 * l2 + c] = isd == 0 ? d : tot[6 * l2 + c] + d;                                 // copy / add2
                }
        }
#pragma unroll
        for (int l2 = 0; l2 < 2; l2++)
#pragma unroll
            for (int c = 0; c < 6; c++) {
                const int ab = lane + 32 * l2;
                if (ab < 36) out[(size_t)e * P2 + c * 36 + ab] = tot[6 * l2 + c];
            }
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
    constexpr size_t smem = (size_t)PN_WARPS * (216 + 36 * FP + 288 + 384) * sizeof(double);
    static bool configured = false;
    if (!configured) {
        NEKB_CUDA(cudaFuncSetAttribute(opgradt_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = true;
    }
    opgradt_kernel<<<grid_for((c.nelv + PN_WARPS - 1) / PN_WARPS, 2), 32 * PN_WARPS, smem, c.stream>>>(ox, oy, oz, p, met9(), mesh2().w3.p,
                                                                                                        c.nelv);
    NEKB_LAUNCHED();
}
inline void opdiv_dev(double *out, const double *ux, const double *uy, const double *uz)
{
    Ctx &c = ctx();
    require_mesh2();
    if (c.nelv <= 0) return;
    constexpr size_t smem = (size_t)PN_WARPS * (64 * UP + 384 + 288) * sizeof(double);
    static bool configured = false;
    if (!configured) {
        NEKB_CUDA(cudaFuncSetAttribute(opdiv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = true;
    }
    opdiv_kernel<<<grid_for((c.nelv + PN_WARPS - 1) / PN_WARPS, 2), 32 * PN_WARPS, smem, c.stream>>>(out, ux, uy, uz, met9(), mesh2().w3.p, c.nelv);
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
