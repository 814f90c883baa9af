// crs_amg_dev.cuh -- device cycle over the host-built aggregation hierarchy of crs_amg.cuh: CG on the assembled coarse
// operator, preconditioned by one V(1,1) cycle (damped Jacobi, CSR transfer operators, dense inverse at the coarsest
// level through the crsd_* kernels of hsmg.cuh).  Plan and iteration counts: DESIGN.md section 8, scripts/proto_coarse_amg.py.
//
// Entry points nekb_crs_amg_upload / nekb_crs_amg_solve_dev; parity test (against numpy on the same levels) in
// tests/test_zz_gpu_configs.py.  Two drivers of the same arithmetic: amg_pcg_solve (one launch per operation, scalars read
// back every iteration: the readable form, NEKB_CRS_AMG_COOP=0) and amg_pcg_coop_kernel (one cooperative launch, default).
#pragma once
#include "crs_amg.cuh"
#include "hsmg.cuh"

namespace nekb {

struct AmgLevelDev {
    int64_t n = 0, nnz = 0;
    DevBuf<int32_t> rowptr, col;       // CSR (nnz < 2^31)
    DevBuf<double> val, dj;            // dj = omega / diag
    DevBuf<int32_t> prow, pcol;        // prolongation P (n x n_next) in CSR: one entry per row (plain aggregation) or the
    DevBuf<double> pval;               // smoothed form
    DevBuf<int32_t> trow, tcol;        // its transpose (n_next x n): restriction
    DevBuf<double> tval;
    DevBuf<double> b, x, x2, r;        // cycle work vectors of this level
};
struct AmgDev {
    bool ready = false;
    std::vector<AmgLevelDev> L;        // every level but the coarsest
    int64_t nc = 0, ld = 0;            // coarsest size, padded leading dimension
    DevBuf<double> ainv, cb, cy;       // explicit inverse, rhs / solution of the coarsest level
    DevBuf<double> r, z, p, p2, w, partial, scal, coop_part;
    DevBuf<int> iters;
};
inline AmgDev &amg_dev()
{
    static AmgDev a;
    return a;
}

constexpr int AMG_T = 256;
inline int amg_grid(int64_t n) { return (int)std::max<int64_t>(1, std::min<int64_t>((n + AMG_T - 1) / AMG_T, (int64_t)ctx().num_sms * 8)); }

// x = dj b (first Jacobi sweep from a zero guess) and r = b - A x, one thread per row
__global__ void __launch_bounds__(AMG_T)
    amg_presmooth_kernel(double *__restrict__ x, double *__restrict__ r, const double *__restrict__ b, const double *__restrict__ dj,
                         const int32_t *__restrict__ rowptr, const int32_t *__restrict__ col, const double *__restrict__ val, int64_t n)
{
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        double s = 0.0;
        for (int q = rowptr[i]; q < rowptr[i + 1]; q++) s = fma(val[q], dj[col[q]] * b[col[q]], s);
        x[i] = dj[i] * b[i];
        r[i] = b[i] - s;
    }
}
// y = A x
__global__ void __launch_bounds__(AMG_T)
    amg_spmv_kernel(double *__restrict__ y, const double *__restrict__ x, const int32_t *__restrict__ rowptr,
                    const int32_t *__restrict__ col, const double *__restrict__ val, int64_t n)
{
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        double s = 0.0;
        for (int q = rowptr[i]; q < rowptr[i + 1]; q++) s = fma(val[q], x[col[q]], s);
        y[i] = s;
    }
}
// xo = x + dj (b - A x)
__global__ void __launch_bounds__(AMG_T)
    amg_postsmooth_kernel(double *__restrict__ xo, const double *__restrict__ x, const double *__restrict__ b,
                          const double *__restrict__ dj, const int32_t *__restrict__ rowptr, const int32_t *__restrict__ col,
                          const double *__restrict__ val, int64_t n)
{
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        double s = 0.0;
        for (int q = rowptr[i]; q < rowptr[i + 1]; q++) s = fma(val[q], x[col[q]], s);
        xo[i] = x[i] + dj[i] * (b[i] - s);
    }
}
// x += P e
__global__ void __launch_bounds__(AMG_T)
    amg_spmv_add_kernel(double *__restrict__ x, const double *__restrict__ e, const int32_t *__restrict__ rowptr,
                        const int32_t *__restrict__ col, const double *__restrict__ val, int64_t n)
{
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        double s = 0.0;
        for (int q = rowptr[i]; q < rowptr[i + 1]; q++) s = fma(val[q], e[col[q]], s);
        x[i] += s;
    }
}
// partial[block] = sum a b over the block's rows; amg_dot_final sums the partials in order into out[slot]
__global__ void __launch_bounds__(AMG_T)
    amg_dot_kernel(double *__restrict__ partial, const double *__restrict__ a, const double *__restrict__ b, int64_t n)
{
    __shared__ double red[33];
    double s = 0.0;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) s = fma(a[i], b[i], s);
    const double t = block_reduce(s, red);
    if (threadIdx.x == 0) partial[blockIdx.x] = t;
}
__global__ void __launch_bounds__(AMG_T) amg_dot_final_kernel(double *__restrict__ out, const double *__restrict__ partial, int nb)
{
    __shared__ double red[33];
    double s = 0.0;
    for (int i = threadIdx.x; i < nb; i += blockDim.x) s += partial[i];
    const double t = block_reduce(s, red);
    if (threadIdx.x == 0) *out = t;
}
// y += a x ; y = x + a y
__global__ void __launch_bounds__(AMG_T) amg_axpy_kernel(double *__restrict__ y, const double *__restrict__ x, double a, int64_t n)
{
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) y[i] = fma(a, x[i], y[i]);
}
__global__ void __launch_bounds__(AMG_T) amg_xpay_kernel(double *__restrict__ y, const double *__restrict__ x, double a, int64_t n)
{
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) y[i] = fma(a, y[i], x[i]);
}

// Moves the host hierarchy to the device and inverts the coarsest operator (blocked Gauss-Jordan of hsmg.cuh).
// fine_mask (may be null): 1 = unmasked dof of the finest level; with null_space the coarsest operator gets gamma m m^T, m = the
// aggregates that hold unmasked dofs (the constant survives every piecewise-constant Galerkin product), exactly as the dense
// path regularises its matrix.
inline void amg_upload(double omega, const std::vector<double> *fine_mask = nullptr, bool null_space = false)
{
    Ctx &c = ctx();
    cudaStream_t s = c.stream;
    const AmgHierarchy &H = amg_host_hierarchy();
    NEKB_REQUIRE(!H.A.empty(), "crs_amg_upload: nekb_crs_amg_build_host has not been called");
    AmgDev &D = amg_dev();
    D = AmgDev();
    const size_t nl = H.A.size() - 1;
    D.L.resize(nl);
    for (size_t l = 0; l < nl; l++) {
        const CsrHost &A = H.A[l];
        AmgLevelDev &L = D.L[l];
        NEKB_REQUIRE(A.nnz() < (int64_t)2147483647, "crs_amg_upload: level too large for int32 offsets");
        L.n = A.n, L.nnz = A.nnz();
        std::vector<int32_t> rp(A.rowptr.begin(), A.rowptr.end());
        std::vector<double> dj((size_t)A.n, 0.0);
        for (int64_t i = 0; i < A.n; i++)
            for (int64_t q = A.rowptr[i]; q < A.rowptr[i + 1]; q++)
                if (A.col[q] == i) dj[i] = omega / A.val[q];
        const CsrHost &P = H.P[l], &PT = H.PT[l];
        NEKB_REQUIRE(P.nnz() < (int64_t)2147483647, "crs_amg_upload: prolongation too large for int32 offsets");
        std::vector<int32_t> pr(P.rowptr.begin(), P.rowptr.end()), tr(PT.rowptr.begin(), PT.rowptr.end());
        L.rowptr.upload(rp.data(), rp.size(), s), L.col.upload(A.col.data(), A.col.size(), s), L.val.upload(A.val.data(), A.val.size(), s);
        L.dj.upload(dj.data(), dj.size(), s);
        L.prow.upload(pr.data(), pr.size(), s), L.pcol.upload(P.col.data(), P.col.size(), s), L.pval.upload(P.val.data(), P.val.size(), s);
        L.trow.upload(tr.data(), tr.size(), s), L.tcol.upload(PT.col.data(), PT.col.size(), s), L.tval.upload(PT.val.data(), PT.val.size(), s);
        L.x.alloc((size_t)A.n), L.x2.alloc((size_t)A.n), L.r.alloc((size_t)A.n);
        if (l > 0) L.b.alloc((size_t)A.n);
        NEKB_CUDA(cudaStreamSynchronize(s));   // the host vectors above go out of scope
    }
    const CsrHost &Ac = H.A.back();
    D.nc = Ac.n, D.ld = crsd_ld(Ac.n);
    std::vector<double> M((size_t)D.ld * D.ld, 0.0);
    for (int64_t i = 0; i < Ac.n; i++)
        for (int64_t q = Ac.rowptr[i]; q < Ac.rowptr[i + 1]; q++) M[(size_t)i * D.ld + Ac.col[q]] = Ac.val[q];
    for (int64_t v = Ac.n; v < D.ld; v++) M[(size_t)v * D.ld + v] = 1.0;
    if (null_space) {
        std::vector<double> m = fine_mask ? *fine_mask : std::vector<double>((size_t)H.A[0].n, 1.0);
        NEKB_REQUIRE((int64_t)m.size() == H.A[0].n, "crs_amg_upload: mask length differs from the finest level");
        for (size_t l = 0; l < nl; l++) {
            std::vector<double> mc((size_t)H.A[l + 1].n, 0.0);
            for (int64_t i = 0; i < H.A[l].n; i++)
                if (m[i] != 0.0) mc[H.agg[l][i]] = 1.0;
            m.swap(mc);
        }
        double trace = 0.0, cnt = 0.0;
        for (int64_t v = 0; v < Ac.n; v++)
            if (m[v] != 0.0) trace += M[(size_t)v * D.ld + v], cnt += 1.0;
        const double gam = cnt > 0.0 ? trace / (cnt * cnt) : 0.0;
        for (int64_t i = 0; i < Ac.n; i++)
            if (m[i] != 0.0)
                for (int64_t j = 0; j < Ac.n; j++)
                    if (m[j] != 0.0) M[(size_t)i * D.ld + j] += gam;
    }
    D.ainv.upload(M.data(), M.size(), s);
    crsd_invert(D.ainv.p, D.ld);
    D.cb.alloc((size_t)D.ld), D.cy.alloc((size_t)D.ld);
    D.cb.zero(s), D.cy.zero(s);
    const int64_t n0 = H.A[0].n;
    D.r.alloc((size_t)n0), D.z.alloc((size_t)n0), D.p.alloc((size_t)n0), D.p2.alloc((size_t)n0), D.w.alloc((size_t)n0);
    D.partial.alloc((size_t)c.num_sms * 8), D.scal.alloc(4);
    NEKB_CUDA(cudaStreamSynchronize(s));
    D.ready = true;
}

// z = M^-1 b: one V(1,1) cycle from level l down; b_dev of level 0 is the caller's, deeper levels use their own buffers
inline void amg_cycle(size_t l, const double *b_dev, double *z_dev)
{
    Ctx &c = ctx();
    cudaStream_t s = c.stream;
    AmgDev &D = amg_dev();
    if (l == D.L.size()) {   // coarsest: z = Ainv b
        const int gw = (int)std::max<int64_t>(1, std::min<int64_t>((D.nc * 32 + 255) / 256, (int64_t)c.num_sms * 8));
        crsd_gemv_kernel<<<gw, 256, 0, s>>>(z_dev, D.ainv.p, b_dev, D.nc, D.ld);
        NEKB_LAUNCHED();
        return;
    }
    AmgLevelDev &L = D.L[l];
    const bool last = l + 1 == D.L.size();
    const int64_t ncn = last ? D.nc : D.L[l + 1].n;
    double *bc = last ? D.cb.p : D.L[l + 1].b.p;
    double *ec = last ? D.cy.p : D.L[l + 1].x2.p;   // the next level leaves its result in its x2 (see the final smoother)
    amg_presmooth_kernel<<<amg_grid(L.n), AMG_T, 0, s>>>(L.x.p, L.r.p, b_dev, L.dj.p, L.rowptr.p, L.col.p, L.val.p, L.n);
    NEKB_LAUNCHED();
    amg_spmv_kernel<<<amg_grid(ncn), AMG_T, 0, s>>>(bc, L.r.p, L.trow.p, L.tcol.p, L.tval.p, ncn);       // bc = P^T r
    NEKB_LAUNCHED();
    amg_cycle(l + 1, bc, ec);
    amg_spmv_add_kernel<<<amg_grid(L.n), AMG_T, 0, s>>>(L.x.p, ec, L.prow.p, L.pcol.p, L.pval.p, L.n);   // x += P e
    NEKB_LAUNCHED();
    amg_postsmooth_kernel<<<amg_grid(L.n), AMG_T, 0, s>>>(z_dev, L.x.p, b_dev, L.dj.p, L.rowptr.p, L.col.p, L.val.p, L.n);
    NEKB_LAUNCHED();
}

inline double amg_dot(const double *a, const double *b, int64_t n)
{
    Ctx &c = ctx();
    AmgDev &D = amg_dev();
    const int nb = std::min(amg_grid(n), (int)D.partial.n);
    amg_dot_kernel<<<nb, AMG_T, 0, c.stream>>>(D.partial.p, a, b, n);
    NEKB_LAUNCHED();
    amg_dot_final_kernel<<<1, AMG_T, 0, c.stream>>>(D.scal.p, D.partial.p, nb);
    NEKB_LAUNCHED();
    double v = 0.0;
    NEKB_CUDA(cudaMemcpyAsync(&v, D.scal.p, sizeof(double), cudaMemcpyDeviceToHost, c.stream));
    NEKB_CUDA(cudaStreamSynchronize(c.stream));
    return v;
}

// CG on A[0] x = b to a relative residual `tol` (2-norm), preconditioned by amg_cycle.  Returns the iteration count.
inline int amg_pcg_solve(double *x_dev, const double *b_dev, double tol, int maxit)
{
    Ctx &c = ctx();
    cudaStream_t s = c.stream;
    AmgDev &D = amg_dev();
    NEKB_REQUIRE(D.ready, "crs_amg_solve: nekb_crs_amg_upload has not been called");
    const bool flat = D.L.empty();   // the whole problem fits the dense inverse
    const int64_t n = flat ? D.nc : D.L[0].n;
    NEKB_CUDA(cudaMemsetAsync(x_dev, 0, sizeof(double) * (size_t)n, s));
    if (flat) {
        amg_cycle(0, b_dev, x_dev);
        return 1;
    }
    AmgLevelDev &L = D.L[0];
    const int g = amg_grid(n);
    NEKB_CUDA(cudaMemcpyAsync(D.r.p, b_dev, sizeof(double) * (size_t)n, cudaMemcpyDeviceToDevice, s));
    const double n0 = std::sqrt(amg_dot(b_dev, b_dev, n));
    if (n0 == 0.0) return 0;
    amg_cycle(0, D.r.p, D.z.p);
    NEKB_CUDA(cudaMemcpyAsync(D.p.p, D.z.p, sizeof(double) * (size_t)n, cudaMemcpyDeviceToDevice, s));
    double rz = amg_dot(D.r.p, D.z.p, n);
    for (int it = 1; it <= maxit; it++) {
        amg_spmv_kernel<<<g, AMG_T, 0, s>>>(D.w.p, D.p.p, L.rowptr.p, L.col.p, L.val.p, n);
        NEKB_LAUNCHED();
        const double a = rz / amg_dot(D.p.p, D.w.p, n);
        amg_axpy_kernel<<<g, AMG_T, 0, s>>>(x_dev, D.p.p, a, n);
        NEKB_LAUNCHED();
        amg_axpy_kernel<<<g, AMG_T, 0, s>>>(D.r.p, D.w.p, -a, n);
        NEKB_LAUNCHED();
        if (std::sqrt(amg_dot(D.r.p, D.r.p, n)) <= tol * n0) return it;
        amg_cycle(0, D.r.p, D.z.p);
        const double rz_new = amg_dot(D.r.p, D.z.p, n);
        amg_xpay_kernel<<<g, AMG_T, 0, s>>>(D.p.p, D.z.p, rz_new / rz, n);   // p = z + beta p
        NEKB_LAUNCHED();
        rz = rz_new;
    }
    return maxit;
}

// ------------------------------------------------------------------------------------------------ one-launch form
// The same CG + V(1,1) cycle as amg_pcg_solve in ONE cooperative launch: the whole hierarchy (<= 10^6 rows) lives in L2, so
// the solve is bound by the number of grid-wide synchronisations, not by bandwidth; the host-driven form above costs ~19
// launches and three scalar read-backs per iteration.  Phases of one iteration (3 levels: 10 grid syncs):
//   w = A (z + beta p) [p rebuilt on the fly, double-buffered]; (p,w)              | sync
//   x += a p; r -= a w; (r,r); level-0 pre-smoothing x0 = dj r, res0 = r - A x0 with the neighbours' new r rebuilt on the fly | sync
//   per level below: restriction | sync | pre-smoothing | sync ... coarsest: y = Ainv b | sync
//   per level upwards: x += P e | sync | post-smoothing (level 0: z, (r,z))        | sync
// Scalars are combined from per-block partials in a fixed order by every block (coop_total), so all blocks take the same
// branch and the result is run-to-run reproducible.
constexpr int AMG_MAXLEV = 6;
struct AmgCoopLevel {
    int n;
    const int32_t *rowptr, *col, *prow, *pcol, *trow, *tcol;
    const double *val, *dj, *pval, *tval;
    double *b, *x, *r, *e;      // rhs (level 0: the CG residual), iterate, residual, result of the level (level 0: z)
};
struct AmgCoopArgs {
    int nlev;                   // levels above the coarsest
    AmgCoopLevel L[AMG_MAXLEV];
    int nc;
    int64_t ld;
    const double *ainv;
    double *cb, *cy;
    double *x, *r, *z, *p0, *p1, *w;
    const double *b;
    double tol;
    int maxit;
    double *partials;           // >= 4 * gridDim doubles
    int *iters_out;
};
// Row sweep of a CSR operator: out(i, sum_q val[q] * xf(col[q])).  Thread per row when there are at least as many rows as
// an eighth of the threads (level 0: 27 entries per row), otherwise warp per row with coalesced row reads (the deeper levels
// and the restrictions, whose rows hold 50-150 entries: a thread walking such a row alone is 150 dependent L2 round trips).
template <class XF, class OUT>
__device__ __forceinline__ void amg_rows(int n, const int32_t *__restrict__ rp, const int32_t *__restrict__ col,
                                         const double *__restrict__ val, int tid, int nth, XF xf, OUT out)
{
    if (n * 8 >= nth) {
        for (int i = tid; i < n; i += nth) {
            double s = 0.0;
            const int e = rp[i + 1];
#pragma unroll 4
            for (int q = rp[i]; q < e; q++) s = fma(val[q], xf(col[q]), s);
            out(i, s);
        }
    } else {
        const int lane = threadIdx.x & 31, nw = nth >> 5;
        for (int i = tid >> 5; i < n; i += nw) {
            double s = 0.0;
            const int e = rp[i + 1];
            for (int q = rp[i] + lane; q < e; q += 32) s = fma(val[q], xf(col[q]), s);
            s = warp_sum(s);
            if (lane == 0) out(i, s);
        }
    }
}

__global__ void __launch_bounds__(1024) amg_pcg_coop_kernel(AmgCoopArgs A)
{
    namespace cgr = cooperative_groups;
    cgr::grid_group grid = cgr::this_grid();
    __shared__ double red[33];
    __shared__ double s_b;
    const int tid = blockIdx.x * blockDim.x + threadIdx.x, nth = gridDim.x * blockDim.x, G = gridDim.x;
    double *P0 = A.partials, *P1 = A.partials + G, *P2 = A.partials + 2 * G, *P3 = A.partials + 3 * G;
    const AmgCoopLevel &F = A.L[0];
    const int n = F.n;
    // x = 0, r = b, (b,b)
    {
        double s = 0.0;
        for (int i = tid; i < n; i += nth) {
            const double bv = A.b[i];
            A.x[i] = 0.0, A.r[i] = bv, A.p0[i] = 0.0;
            s = fma(bv, bv, s);
        }
        const double t = block_reduce(s, red);
        if (threadIdx.x == 0) P0[blockIdx.x] = t;
    }
    grid.sync();
    const double bb = coop_total(P0, G, red, &s_b);
    int it = 0;
    if (bb > 0.0) {
        const double stop2 = A.tol * A.tol * bb;
        // one V(1,1) cycle on the current r -> z, with (r,z) partials in P1.  `fresh` = the level-0 pre-smoothing has already
        // been done by the caller phase (fused with the r update).
        auto cycle = [&](bool fresh) {
            for (int l = 0; l < A.nlev; l++) {
                const AmgCoopLevel &L = A.L[l];
                if (!(l == 0 && fresh)) {
                    const double *bl = l == 0 ? A.r : L.b;
                    amg_rows(L.n, L.rowptr, L.col, L.val, tid, nth, [&](int jc) { return L.dj[jc] * bl[jc]; },
                             [&](int i, double sm) {            // x = dj b ; r = b - A x
                                 L.x[i] = L.dj[i] * bl[i];
                                 L.r[i] = bl[i] - sm;
                             });
                    grid.sync();
                }
                const bool last = l + 1 == A.nlev;
                double *bc = last ? A.cb : A.L[l + 1].b;
                const int ncn = last ? A.nc : A.L[l + 1].n;
                amg_rows(ncn, L.trow, L.tcol, L.tval, tid, nth, [&](int jc) { return L.r[jc]; },
                         [&](int i, double sm) { bc[i] = sm; });     // bc = P^T r
                grid.sync();
            }
            {   // coarsest: y = Ainv b, one warp per row
                const int lane = threadIdx.x & 31, w0 = tid >> 5, nw = nth >> 5;
                for (int r = w0; r < A.nc; r += nw) {
                    const double *row = A.ainv + (size_t)r * A.ld;
                    double s = 0.0;
                    for (int j = lane; j < A.nc; j += 32) s = fma(row[j], A.cb[j], s);
                    s = warp_sum(s);
                    if (lane == 0) A.cy[r] = s;
                }
            }
            grid.sync();
            for (int l = A.nlev - 1; l >= 0; l--) {
                const AmgCoopLevel &L = A.L[l];
                const bool last = l + 1 == A.nlev;
                const double *ec = last ? A.cy : A.L[l + 1].e;
                amg_rows(L.n, L.prow, L.pcol, L.pval, tid, nth, [&](int jc) { return ec[jc]; },
                         [&](int i, double sm) { L.x[i] += sm; });   // x += P e
                grid.sync();
                const double *bl = l == 0 ? A.r : L.b;
                double *out = l == 0 ? A.z : L.e;
                double rzp = 0.0;
                amg_rows(L.n, L.rowptr, L.col, L.val, tid, nth, [&](int jc) { return L.x[jc]; },
                         [&](int i, double sm) {                  // out = x + dj (b - A x)
                             const double v = L.x[i] + L.dj[i] * (bl[i] - sm);
                             out[i] = v;
                             rzp = fma(bl[i], v, rzp);
                         });
                if (l == 0) {
                    const double t = block_reduce(rzp, red);
                    if (threadIdx.x == 0) P1[blockIdx.x] = t;
                }
                grid.sync();
            }
        };
        cycle(false);
        double rz = coop_total(P1, G, red, &s_b);
        double beta = 0.0;
        double *pin = A.p0, *pout = A.p1;
        for (; it < A.maxit;) {
            {   // p = z + beta p (own entry stored, neighbours rebuilt) ; w = A p ; (p,w)
                double s = 0.0;
                amg_rows(n, F.rowptr, F.col, F.val, tid, nth, [&](int jc) { return fma(beta, pin[jc], A.z[jc]); },
                         [&](int i, double acc) {
                             const double pi = fma(beta, pin[i], A.z[i]);
                             pout[i] = pi;
                             A.w[i] = acc;
                             s = fma(pi, acc, s);
                         });
                const double t = block_reduce(s, red);
                if (threadIdx.x == 0) P2[blockIdx.x] = t;
            }
            grid.sync();
            const double pw = coop_total(P2, G, red, &s_b);
            const double alpha = rz / pw;
            {   // x += a p ; r -= a w ; (r,r) ; level-0 pre-smoothing on the new r (neighbours' r rebuilt: r_j - a w_j)
                double s = 0.0;
                amg_rows(n, F.rowptr, F.col, F.val, tid, nth, [&](int jc) { return F.dj[jc] * fma(-alpha, A.w[jc], A.r[jc]); },
                         [&](int i, double acc) {
                             A.x[i] = fma(alpha, pout[i], A.x[i]);
                             const double ri = fma(-alpha, A.w[i], A.r[i]);
                             F.x[i] = F.dj[i] * ri;
                             F.r[i] = ri - acc;
                             A.z[i] = ri;          // parked: A.r is still being read by other threads' neighbour sums
                             s = fma(ri, ri, s);
                         });
                const double t = block_reduce(s, red);
                if (threadIdx.x == 0) P3[blockIdx.x] = t;
            }
            grid.sync();
            // own entries only (the writer of z[i] in the sweep above is the thread that copies it here when rows are swept by
            // threads; with the warp-per-row sweep lane 0 wrote it, so every thread copies a strided share after the sync)
            for (int i = tid; i < n; i += nth) A.r[i] = A.z[i];
            const double rr = coop_total(P3, G, red, &s_b);
            it++;
            if (rr <= stop2) break;
            cycle(true);   // reads r at other rows only in its last phase (level-0 post-smoothing), several syncs from here
            const double rz_new = coop_total(P1, G, red, &s_b);
            beta = rz_new / rz;
            rz = rz_new;
            double *tmp = pin;
            pin = pout, pout = tmp;
        }
    }
    if (tid == 0) *A.iters_out = it;
}

inline int amg_coop_enabled()
{
    const char *e = getenv("NEKB_CRS_AMG_COOP");
    return e ? atoi(e) : 1;
}

// Launches the one-kernel solve; the iteration count stays on the device (D.iters) unless `want_iters`.
inline int amg_pcg_solve_coop(double *x_dev, const double *b_dev, double tol, int maxit, bool want_iters)
{
    Ctx &c = ctx();
    cudaStream_t s = c.stream;
    AmgDev &D = amg_dev();
    NEKB_REQUIRE(D.ready && !D.L.empty() && D.L.size() <= (size_t)AMG_MAXLEV, "amg_pcg_solve_coop: hierarchy not uploaded / too deep");
    AmgCoopArgs A;
    memset(&A, 0, sizeof A);
    A.nlev = (int)D.L.size();
    for (int l = 0; l < A.nlev; l++) {
        AmgLevelDev &L = D.L[l];
        AmgCoopLevel &o = A.L[l];
        o.n = (int)L.n;
        o.rowptr = L.rowptr.p, o.col = L.col.p, o.val = L.val.p, o.dj = L.dj.p;
        o.prow = L.prow.p, o.pcol = L.pcol.p, o.pval = L.pval.p, o.trow = L.trow.p, o.tcol = L.tcol.p, o.tval = L.tval.p;
        o.b = l > 0 ? L.b.p : nullptr, o.x = L.x.p, o.r = L.r.p, o.e = L.x2.p;
    }
    A.nc = (int)D.nc, A.ld = D.ld, A.ainv = D.ainv.p, A.cb = D.cb.p, A.cy = D.cy.p;
    A.x = x_dev, A.r = D.r.p, A.z = D.z.p, A.p0 = D.p.p, A.p1 = D.p2.p, A.w = D.w.p, A.b = b_dev;
    A.tol = tol, A.maxit = maxit;
    static int per_sm = 0;
    if (!per_sm) {
        NEKB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, amg_pcg_coop_kernel, 1024, 0));
        NEKB_REQUIRE(per_sm >= 1, "amg_pcg_coop_kernel cannot be made resident");
        per_sm = 1;   // one CTA per SM: the phases are synchronisation-bound, fewer arrivals make a cheaper grid sync
    }
    int gridc = (int)((D.L[0].n + 1023) / 1024);
    if (gridc > c.num_sms * per_sm) gridc = c.num_sms * per_sm;
    D.coop_part.ensure((size_t)4 * c.num_sms * per_sm);
    D.iters.ensure(1);
    A.partials = D.coop_part.p, A.iters_out = D.iters.p;
    void *args[] = {&A};
    NEKB_CUDA(cudaLaunchCooperativeKernel((void *)amg_pcg_coop_kernel, dim3(gridc), dim3(1024), args, 0, s));
    NEKB_LAUNCHED();
    if (!want_iters) return -1;
    int it = 0;
    NEKB_CUDA(cudaMemcpyAsync(&it, D.iters.p, sizeof(int), cudaMemcpyDeviceToHost, s));
    NEKB_CUDA(cudaStreamSynchronize(s));
    return it;
}

// ------------------------------------------------------------------------------------------------ h1mg wiring
// crs_solve semantics (see crs_dense_solve) with the explicit inverse replaced by amg_pcg_solve on the assembled operator:
// gather to the distinct dofs, all-reduce, mask / mean removal, CG to k.tol, mean removal, scatter.  Every rank holds the
// whole hierarchy and solves redundantly, so the CG itself needs no communication.
inline void crs_amg_solve(CrsSolver &k, double *x_out, const double *b_in)
{
    Ctx &c = ctx();
    cudaStream_t s = c.stream;
    const int64_t nc = k.nc;
    const int gv = (int)std::max<int64_t>(1, std::min<int64_t>((nc + 255) / 256, (int64_t)c.num_sms * 4));
    crsd_gather_kernel<<<gv, 256, 0, s>>>(k.g.p, b_in, k.voff.p, k.vmem.p, nc);
    NEKB_LAUNCHED();
    if (c.nranks > 1) comm_allreduce_sum(k.g.p, (int)nc);
    crsd_prep_kernel<<<1, 1024, 0, s>>>(k.g.p, k.gmask.p, nc, k.null_space, k.ndof);
    NEKB_LAUNCHED();
    if (k.amg_one_launch) {                                             // decided at set-up (NEKB_CRS_AMG_COOP)
        amg_pcg_solve_coop(k.y.p, k.g.p, k.tol, k.maxit, false);      // no host round trip inside h1mg_solve
        k.last_iters = -1, k.iters_on_device = false;
    } else {
        k.last_iters = amg_pcg_solve(k.y.p, k.g.p, k.tol, k.maxit);
        k.iters_on_device = false;
    }
    crsd_mean_kernel<<<1, 1024, 0, s>>>(k.y.p, k.gmask.p, nc, k.null_space, k.ndof);
    NEKB_LAUNCHED();
    if (k.n > 0) {
        const int gx = (int)std::max<int64_t>(1, std::min<int64_t>((k.n + 255) / 256, (int64_t)c.num_sms * 4));
        crsd_scatter_kernel<<<gx, 256, 0, s>>>(x_out, k.y.p, k.vid.p, k.n);
        NEKB_LAUNCHED();
    }
}

// Same gathering of every rank's ids / element matrices / masks as crs_dense_setup, assembled into COO triplets (masked dofs:
// identity rows) instead of a dense matrix.  COLLECTIVE.  Only with NEKB_CRS_AMG=1.
inline void crs_amg_setup(CrsSolver &k, int nel, const int64_t *vertex)
{
    const char *env = getenv("NEKB_CRS_AMG");
    if (!env || atoi(env) == 0) return;
    Ctx &c = ctx();
    cudaStream_t s = c.stream;
    std::vector<int64_t> glo((size_t)8 * std::max(nel, 1));
    setvert3d_host(glo.data(), 2, nel, vertex, c.nranks);
    int64_t mx = 0;
    for (int t = 0; t < 8 * nel; t++) mx = std::max(mx, glo[t]);
    std::vector<int64_t> all((size_t)c.nranks);
    host_allgather(&mx, all.data(), sizeof(int64_t));
    int64_t nc = 0;
    for (int64_t v : all) nc = std::max(nc, v);
    if (nc <= 0) return;
    int64_t nel_loc = nel;
    std::vector<int64_t> nels((size_t)c.nranks);
    host_allgather(&nel_loc, nels.data(), sizeof(int64_t));
    int64_t nelmax = 1;
    for (int64_t v : nels) nelmax = std::max(nelmax, v);
    std::vector<double> a_loc((size_t)64 * nelmax, 0.0), m_loc((size_t)8 * nelmax, 0.0);
    std::vector<int64_t> id_loc((size_t)8 * nelmax, 0);
    if (nel > 0) {
        k.a.download(a_loc.data(), (size_t)64 * nel, s);
        k.mask.download(m_loc.data(), (size_t)8 * nel, s);
        memcpy(id_loc.data(), glo.data(), sizeof(int64_t) * 8 * (size_t)nel);
    }
    std::vector<double> a_all((size_t)64 * nelmax * c.nranks), m_all((size_t)8 * nelmax * c.nranks);
    std::vector<int64_t> id_all((size_t)8 * nelmax * c.nranks);
    host_allgather(a_loc.data(), a_all.data(), sizeof(double) * a_loc.size());
    host_allgather(m_loc.data(), m_all.data(), sizeof(double) * m_loc.size());
    host_allgather(id_loc.data(), id_all.data(), sizeof(int64_t) * id_loc.size());
    std::vector<double> gmask((size_t)nc, 1.0);
    for (int r = 0; r < c.nranks; r++)
        for (int64_t e = 0; e < nels[r]; e++)
            for (int i = 0; i < 8; i++) {
                const int64_t g = id_all[((size_t)r * nelmax + e) * 8 + i];
                NEKB_REQUIRE(g >= 1 && g <= nc, "coarse numbering out of range");
                if (m_all[((size_t)r * nelmax + e) * 8 + i] == 0.0) gmask[g - 1] = 0.0;
            }
    std::vector<int64_t> I, J;
    std::vector<double> V;
    for (int r = 0; r < c.nranks; r++)
        for (int64_t e = 0; e < nels[r]; e++) {
            const double *ae = a_all.data() + ((size_t)r * nelmax + e) * 64;
            const int64_t *ie = id_all.data() + ((size_t)r * nelmax + e) * 8;
            for (int i = 0; i < 8; i++)
                for (int j = 0; j < 8; j++)
                    if (gmask[ie[i] - 1] != 0.0 && gmask[ie[j] - 1] != 0.0) I.push_back(ie[i] - 1), J.push_back(ie[j] - 1), V.push_back(ae[i * 8 + j]);
        }
    for (int64_t v = 0; v < nc; v++)
        if (gmask[v] == 0.0) I.push_back(v), J.push_back(v), V.push_back(1.0);
    const char *en = getenv("NEKB_CRS_AMG_NMAX");
    const int64_t nmax = en ? atoll(en) : 2048;
    const char *es = getenv("NEKB_CRS_AMG_SMOOTH");   // prolongation smoothing weight; 0 = plain aggregation
    const double omega_p = es ? atof(es) : 0.66;
    amg_host_hierarchy() = amg_build(csr_from_triplets(nc, I, J, V), nmax, 0.02, omega_p);
    amg_upload(0.7, &gmask, k.null_space != 0);
    // local members of every global dof, as in crs_dense_setup
    std::vector<int32_t> vid((size_t)8 * std::max(nel, 1), 0), voff((size_t)nc + 1, 0), vmem((size_t)8 * std::max(nel, 1), 0);
    for (int t = 0; t < 8 * nel; t++) vid[t] = (int32_t)(glo[t] - 1), voff[(size_t)vid[t] + 1]++;
    for (int64_t v = 0; v < nc; v++) voff[v + 1] += voff[v];
    std::vector<int32_t> cur(voff.begin(), voff.end() - 1);
    for (int t = 0; t < 8 * nel; t++) vmem[cur[vid[t]]++] = t;
    k.vid.upload(vid.data(), vid.size(), s), k.voff.upload(voff.data(), voff.size(), s), k.vmem.upload(vmem.data(), vmem.size(), s);
    k.gmask.upload(gmask.data(), (size_t)nc, s);
    k.g.alloc((size_t)nc), k.y.alloc((size_t)nc);
    k.g.zero(s), k.y.zero(s);
    k.nc = nc;
    NEKB_CUDA(cudaStreamSynchronize(s));
    k.amg_solve = crs_amg_solve;
    k.amg_one_launch = amg_coop_enabled() && !amg_dev().L.empty();
}

struct CrsAmgHookInstaller {
    CrsAmgHookInstaller() { crs_amg_setup_hook() = crs_amg_setup; }
};
static CrsAmgHookInstaller crs_amg_hook_installer;

}  // namespace nekb
