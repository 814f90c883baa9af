// proj.cuh -- residual projection around the Helmholtz / pressure solves: hsolve (core/navier4.f:562-634), project1 (:636-733),
// project1_a (:735-793), iproj_chk (:795-826), proj_matvec (:828-845), proj_ortho (:848-945), proj_ortho_full_cgs2
// (:1060-1125), project2 / project2_a (:1129-1199), hmhzpf (:513-560).  SURVEY.md 8f rank 2.
//
// The approximation space X (previous solutions, A-orthonormal), B = A X, xbar, bbar and the saved h1/h2 live on the device,
// one set per solver name (the reference keeps them in the caller's `approx` array; only ivar(1:2) = mmx, m are mirrored
// back).  The few scalars of a projection (at most mmx = 8 coefficients per Gram-Schmidt round) are reduced on the device
// and read back once per round -- this runs once per solve, not per iteration.
#pragma once
#include <map>
#include <string>

#include "hcg.cuh"

namespace nekb {

constexpr int PROJ_MAXP = 16;  // dot-product pairs per launch (2 * mmx)
struct ProjPairs {
    const double *a[PROJ_MAXP], *c[PROJ_MAXP];
};

// out[k] = sum_i a_k[i] * w[i] * c_k[i]  (vlsc3; w == nullptr: vlsc2), k < np, one pass over w
__global__ void __launch_bounds__(256)
    proj_dots_kernel(ProjPairs P, const double *__restrict__ w, int np, int64_t n, double *out, double *partials, unsigned *counter)
{
    __shared__ double red[33];
    double s[PROJ_MAXP];
#pragma unroll
    for (int k = 0; k < PROJ_MAXP; k++) s[k] = 0.0;
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < n; t += (int64_t)gridDim.x * blockDim.x) {
        const double wt = w ? w[t] : 1.0;
#pragma unroll
        for (int k = 0; k < PROJ_MAXP; k++)
            if (k < np) s[k] = fma(P.a[k][t] * wt, P.c[k][t], s[k]);
    }
    for (int k = 0; k < np; k++) {
        const double b = block_reduce(s[k], red);
        double *o = out + k;
        grid_reduce(b, partials + (size_t)k * gridDim.x, counter + k, red, [=](double tot) { *o = tot; });
    }
}
__global__ void __launch_bounds__(256) proj_axpy_kernel(double *__restrict__ y, const double *__restrict__ x, double a, int64_t n)
{
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < n; t += (int64_t)gridDim.x * blockDim.x) y[t] = fma(a, x[t], y[t]);
}
__global__ void __launch_bounds__(256) proj_givens_kernel(double *__restrict__ h, double *__restrict__ k, double c, double s, int64_t n)
{
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < n; t += (int64_t)gridDim.x * blockDim.x) {
        const double a = h[t], b = k[t];
        h[t] = c * a + s * b;
        k[t] = -s * a + c * b;
    }
}
__global__ void __launch_bounds__(256)
    proj_maxdiff_kernel(const double *__restrict__ a, const double *__restrict__ a0, const double *__restrict__ b, const double *__restrict__ b0,
                        int64_t n, double *out, double *partials, unsigned *counter)
{
    __shared__ double red[33];
    double m = 0.0;
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < n; t += (int64_t)gridDim.x * blockDim.x)
        m = fmax(m, fmax(fabs(a[t] - a0[t]), fabs(b[t] - b0[t])));
    const double r = block_reduce<true>(m, red);
    grid_reduce<true>(r, partials, counter, red, [=](double tot) { *out = tot; });
}

struct ProjState {
    int64_t n = 0;
    int m = 0, mmx = 0;
    std::vector<DevBuf<double>> X, B;
    DevBuf<double> xbar, bbar, h1old, h2old;
};
inline std::map<std::string, ProjState> &proj_states()
{
    static std::map<std::string, ProjState> s;
    return s;
}
struct ProjScratch {
    DevBuf<double> out, partials;
    DevBuf<unsigned> counters;   // one grid_reduce ticket per dot product of a launch
    void ensure(cudaStream_t s)
    {
        out.ensure(PROJ_MAXP), partials.ensure((size_t)PROJ_MAXP * CG_PART_STRIDE);
        if (!counters.p) {
            counters.alloc(PROJ_MAXP);
            counters.zero(s);
        }
    }
};
inline ProjScratch &proj_scratch()
{
    static ProjScratch s;
    return s;
}

inline void proj_dots(const ProjPairs &P, int np, const double *w, int64_t n, double *host)
{
    Ctx &c = ctx();
    ProjScratch &S = proj_scratch();
    const int grid = cg_grid(n);
    S.ensure(c.stream);
    proj_dots_kernel<<<grid, 256, 0, c.stream>>>(P, w, np, n, S.out.p, S.partials.p, S.counters.p);
    NEKB_LAUNCHED();
    comm_allreduce_sum(S.out.p, np);   // gop(alpha,work,'+  ',m)
    NEKB_CUDA(cudaMemcpyAsync(host, S.out.p, sizeof(double) * np, cudaMemcpyDeviceToHost, c.stream));
    NEKB_CUDA(cudaStreamSynchronize(c.stream));
}
inline void proj_axpy(double *y, const double *x, double a, int64_t n)
{
    proj_axpy_kernel<<<cg_grid(n), 256, 0, ctx().stream>>>(y, x, a, n);
    NEKB_LAUNCHED();
}
inline void proj_scale(double *x, double a, int64_t n)
{
    gm_cmult2_kernel<<<cg_grid(n), 256, 0, ctx().stream>>>(x, x, a, n);
    NEKB_LAUNCHED();
}
inline void proj_copy(double *dst, const double *src, int64_t n)
{
    NEKB_CUDA(cudaMemcpyAsync(dst, src, sizeof(double) * (size_t)n, cudaMemcpyDeviceToDevice, ctx().stream));
}
// navier4.f:828-845 proj_matvec: b = mask * dssum(A x)
inline void proj_matvec(double *b, const double *x, const double *h1, const double *h2, const double *msk, int nel, int gs_handle, bool ifh2)
{
    launch_ax(x, b, h1, ifh2 ? h2 : nullptr, nel, nullptr);
    gs_op(gs_handle, b, 1, msk);
}
// alpha(j) = .5*( (xx_j, w bb_k) + (bb_j, w xx_k) ) for j in [j0, j1]
inline void proj_sym_dots(ProjState &S, int j0, int j1, int k, const double *w, double *alpha)
{
    ProjPairs P;
    int np = 0;
    for (int j = j0; j <= j1; j++) {
        P.a[np] = S.X[j].p, P.c[np] = S.B[k].p, np++;
        P.a[np] = S.B[j].p, P.c[np] = S.X[k].p, np++;
    }
    if (np == 0) return;
    double h[PROJ_MAXP];
    proj_dots(P, np, w, S.n, h);
    for (int j = j0; j <= j1; j++) alpha[j] = 0.5 * (h[2 * (j - j0)] + h[2 * (j - j0) + 1]);
}

// navier4.f:1060-1125 proj_ortho_full_cgs2 (0-based vector indices)
inline void proj_ortho_full_cgs2(ProjState &S, const double *w)
{
    int m = S.m;
    if (m <= 0) return;
    const double tol = 1.e-7;
    std::vector<int> flag(m, 0);
    std::vector<double> alpha(m, 0.0);
    for (int pass = 0; pass < 2; pass++)
        for (int k = m - 1; k >= 0; k--) {
            proj_sym_dots(S, k, m - 1, k, w, alpha.data());
            for (int j = m - 1; j >= k + 1; j--) {
                proj_axpy(S.X[k].p, S.X[j].p, -alpha[j], S.n);
                proj_axpy(S.B[k].p, S.B[j].p, -alpha[j], S.n);
            }
            const double normp = sqrt(alpha[k]);
            ProjPairs P;
            P.a[0] = S.X[k].p, P.c[0] = S.B[k].p;
            double nk = 0.0;
            proj_dots(P, 1, w, S.n, &nk);
            const double normk = sqrt(nk);
            if (normk > tol * normp) {
                proj_scale(S.X[k].p, 1.0 / normk, S.n);
                proj_scale(S.B[k].p, 1.0 / normk, S.n);
                flag[k] = 1;
            } else
                flag[k] = 0;
        }
    int k = 0;
    for (int j = 0; j < m; j++)
        if (flag[j]) {
            if (k < j) std::swap(S.X[k], S.X[j]), std::swap(S.B[k], S.B[j]);
            k++;
        }
    S.m = k;
}

// core/navier4.f:1308-1344 givens_rotation(a,b,c,s,r) with the reference's own hypot
inline void givens_rotation(double a, double b, double &c, double &s, double &r)
{
    if (b != 0.0) {
        const double ca = fabs(a), cb = fabs(b), x = ca > cb ? ca : cb, t = (1.0 / x) * (ca < cb ? ca : cb);
        const double h = x * sqrt(1.0 + t * t), d = 1.0 / h;
        c = fabs(a) * d;
        s = copysign(d, a) * b;
        r = copysign(1.0, a) * h;
    } else {
        c = 1.0, s = 0.0, r = a;
    }
}

// navier4.f:848-945 proj_ortho: the newest vector (index m-1) is A-orthonormalised against the others (CGS2) and the basis is
// rotated so that the oldest information leaves first
inline void proj_ortho(ProjState &S, const double *w)
{
    const int m = S.m;
    if (m <= 0) return;
    std::vector<double> alpha(m, 0.0), beta(m, 0.0);
    proj_sym_dots(S, 0, m - 1, m - 1, w, alpha.data());
    double nrm = sqrt(alpha[m - 1]);
    for (int k = 0; k < m - 1; k++) {
        proj_axpy(S.X[m - 1].p, S.X[k].p, -alpha[k], S.n);
        proj_axpy(S.B[m - 1].p, S.B[k].p, -alpha[k], S.n);
    }
    if (m > 1) proj_sym_dots(S, 0, m - 2, m - 1, w, beta.data());
    for (int k = 0; k < m - 1; k++) {
        proj_axpy(S.X[m - 1].p, S.X[k].p, -beta[k], S.n);
        proj_axpy(S.B[m - 1].p, S.B[k].p, -beta[k], S.n);
        alpha[k] = alpha[k] + beta[k];
    }
    ProjPairs P;
    P.a[0] = S.X[m - 1].p, P.c[0] = S.B[m - 1].p;
    double am = 0.0;
    proj_dots(P, 1, w, S.n, &am);
    alpha[m - 1] = sqrt(am);
    const double tol = 1.e-7;
    if (alpha[m - 1] > tol * nrm) {
        const double scl = 1.0 / alpha[m - 1];
        proj_scale(S.X[m - 1].p, scl, S.n);
        proj_scale(S.B[m - 1].p, scl, S.n);
        for (int k = m - 1; k >= 1; k--) {
            const int h = k - 1;
            double c, s;
            givens_rotation(alpha[h], alpha[k], c, s, nrm);
            alpha[h] = nrm;
            proj_givens_kernel<<<cg_grid(S.n), 256, 0, ctx().stream>>>(S.X[h].p, S.X[k].p, c, s, S.n);
            NEKB_LAUNCHED();
            proj_givens_kernel<<<cg_grid(S.n), 256, 0, ctx().stream>>>(S.B[h].p, S.B[k].p, c, s, S.n);
            NEKB_LAUNCHED();
        }
    } else
        S.m = m - 1;  // rank deficient: forget the new vector
}

inline ProjState &proj_get(const std::string &name6, int64_t n, int mxprev)
{
    ProjState &S = proj_states()[name6];
    const int mmx = (mxprev - 4) / 2;  // proj_get_ivar, navier4.f:1216
    NEKB_REQUIRE(mmx >= 1 && 2 * mmx <= PROJ_MAXP, "projection: mxprev out of range (6 <= mxprev <= 20)");
    if (S.n != n || S.mmx != mmx) {
        S = ProjState();
        S.n = n, S.mmx = mmx;
        S.X.resize(mmx), S.B.resize(mmx);
        for (int k = 0; k < mmx; k++) S.X[k].alloc((size_t)n), S.B[k].alloc((size_t)n);
        S.xbar.alloc((size_t)n), S.bbar.alloc((size_t)n), S.h1old.alloc((size_t)n), S.h2old.alloc((size_t)n);
        S.h1old.zero(ctx().stream), S.h2old.zero(ctx().stream);
    }
    return S;
}

// navier4.f:636-733 project1 (+ project1_a): b <- b - B X^T b ; xbar = X X^T b.  Returns bb4/baf for the log line.
inline double project1_dev(ProjState &S, double *b, const double *h1, const double *h2, const double *msk, const double *w, int nel,
                           int gs_handle, bool ifh2)
{
    Ctx &c = ctx();
    if (S.m <= 0) return 1.0;
    const int64_t n = S.n;
    // iproj_chk (:795-826): the matrix changed if h1 / h2 differ from the saved copies
    ProjScratch &R = proj_scratch();
    R.ensure(c.stream);
    proj_maxdiff_kernel<<<cg_grid(n), 256, 0, c.stream>>>(h1, S.h1old.p, h2, S.h2old.p, n, R.out.p, R.partials.p, R.counters.p);
    NEKB_LAUNCHED();
    comm_allreduce_max(R.out.p, 1);
    double dh = 0.0;
    NEKB_CUDA(cudaMemcpyAsync(&dh, R.out.p, sizeof(double), cudaMemcpyDeviceToHost, c.stream));
    NEKB_CUDA(cudaStreamSynchronize(c.stream));
    ProjPairs P;
    P.a[0] = b, P.c[0] = b;
    double bb4 = 0.0;
    proj_dots(P, 1, w, n, &bb4);
    bb4 = sqrt(bb4);
    if (dh > 0.0) {
        proj_copy(S.h1old.p, h1, n), proj_copy(S.h2old.p, h2, n);
        for (int j = 0; j < S.m; j++) proj_matvec(S.B[j].p, S.X[j].p, h1, h2, msk, nel, gs_handle, ifh2);
        proj_ortho_full_cgs2(S, w);
        if (S.m <= 0) return 1.0;
    }
    // project1_a: two rounds of classical Gram-Schmidt
    const int m = S.m;
    double alpha[PROJ_MAXP];
    for (int round = 0; round < 2; round++) {
        for (int k = 0; k < m; k++) P.a[k] = S.X[k].p, P.c[k] = b;
        proj_dots(P, m, w, n, alpha);
        for (int k = 0; k < m; k++) {
            if (round == 0 && k == 0) {
                gm_cmult2_kernel<<<cg_grid(n), 256, 0, c.stream>>>(S.xbar.p, S.X[0].p, alpha[0], n);
                NEKB_LAUNCHED();
                gm_cmult2_kernel<<<cg_grid(n), 256, 0, c.stream>>>(S.bbar.p, S.B[0].p, alpha[0], n);
                NEKB_LAUNCHED();
            } else {
                proj_axpy(S.xbar.p, S.X[k].p, alpha[k], n);
                proj_axpy(S.bbar.p, S.B[k].p, alpha[k], n);
            }
            proj_axpy(b, S.B[k].p, -alpha[k], n);
        }
    }
    P.a[0] = b, P.c[0] = b;
    double baf = 0.0;
    proj_dots(P, 1, w, n, &baf);
    baf = sqrt(baf);
    return baf > 0.0 ? bb4 / baf : 0.0;
}

// navier4.f:1129-1199 project2 (+ project2_a): x += xbar ; push x into the space and re-orthogonalise
inline void project2_dev(ProjState &S, double *x, const double *h1, const double *h2, const double *msk, const double *w, int nel,
                         int gs_handle, bool ifh2)
{
    const int64_t n = S.n;
    if (S.m > 0) proj_axpy(x, S.xbar.p, 1.0, n);
    S.m = S.m + 1 < S.mmx ? S.m + 1 : S.mmx;
    proj_copy(S.X[S.m - 1].p, x, n);
    proj_matvec(S.B[S.m - 1].p, S.X[S.m - 1].p, h1, h2, msk, nel, gs_handle, ifh2);
    proj_ortho(S, w);
}

}  // namespace nekb
