// setup.cuh -- everything that feeds the hot path: GLL operators, global node numbering (host, integer,
// bit-exact target), geometric factors (device), and the synthetic BP5 case.
//
// Reference: core/speclib.f:107 (ZWGLL), :800 (DGLL); core/navier8.f:2004-2360 (setvert3d) with
// :1934-2002 (gbtuple_rank8); core/genxyz.f:1269-1332 (xyzlin); core/coef.f:555-784 (glmapm1, geodat1);
// examples/bp5/bp5.usr:623-699 (geodatstd), :142-153 (xmask1), :324-395 (bp5); core/navier5.f:2650-2700
// (ran1, rand_fld_h1); core/ic.f:1871-1895 (dsavg); core/connect1.f:124-135 (vmult).
#pragma once
#include <cmath>

#include "cg.cuh"
#include "comm.cuh"

namespace nekb {

// ------------------------------------------------------------------------------------------------ GLL operators
// Used only when the host program has not registered its own zgm1/wxm1/dxm1 (nekb_set_gll/nekb_set_dxyz).
// Newton iteration on (1-x^2) P_N'(x) in long double; D from the Lagrange formula on the GLL nodes.
inline void legendre(int N, long double x, long double &p, long double &dp)
{
    long double p0 = 1.0L, p1 = x;
    if (N == 0) {
        p = 1.0L, dp = 0.0L;
        return;
    }
    for (int k = 2; k <= N; k++) {
        long double pk = ((2 * k - 1) * x * p1 - (k - 1) * p0) / k;
        p0 = p1;
        p1 = pk;
    }
    p = p1;
    dp = (x * x == 1.0L) ? (x > 0 ? 1.0L : ((N % 2) ? 1.0L : -1.0L)) * N * (N + 1) / 2.0L
                         : N * (p0 - x * p1) / (1.0L - x * x);
}

inline void gll_build(int nx, std::vector<double> &z, std::vector<double> &w, std::vector<double> &D)
{
    const int N = nx - 1;
    std::vector<long double> zz(nx);
    const long double pi = 3.141592653589793238462643383279502884L;
    zz[0] = -1.0L;
    zz[N] = 1.0L;
    for (int i = 1; i < N; i++) {
        long double x = -cosl(pi * i / N);
        for (int it = 0; it < 100; it++) {
            long double p, dp;
            legendre(N, x, p, dp);
            // f = (1-x^2) P'  ;  f' = -2x P' + (1-x^2) P'' = -N(N+1) P   (Legendre ODE)
            long double f = (1.0L - x * x) * dp, fp = -(long double)N * (N + 1) * p;
            long double dx = f / fp;
            x -= dx;
            if (fabsl(dx) < 1e-19L) break;
        }
        zz[i] = x;
    }
    for (int i = 0; i <= N / 2; i++) {  // enforce symmetry
        long double a = 0.5L * (zz[N - i] - zz[i]);
        zz[i] = -a, zz[N - i] = a;
    }
    if (N % 2 == 0) zz[N / 2] = 0.0L;
    z.resize(nx), w.resize(nx), D.assign((size_t)nx * nx, 0.0);
    std::vector<long double> pn(nx);
    for (int i = 0; i < nx; i++) {
        long double p, dp;
        legendre(N, zz[i], p, dp);
        pn[i] = p;
        z[i] = (double)zz[i];
        w[i] = (double)(2.0L / ((long double)N * (N + 1) * p * p));
    }
    for (int i = 0; i < nx; i++)
        for (int j = 0; j < nx; j++) {
            long double d;
            if (i != j)
                d = pn[i] / (pn[j] * (zz[i] - zz[j]));
            else if (i == 0)
                d = -(long double)N * (N + 1) / 4.0L;
            else if (i == N)
                d = (long double)N * (N + 1) / 4.0L;
            else
                d = 0.0L;
            D[(size_t)i * nx + j] = (double)d;
        }
}

inline void ensure_operators()
{
    Ctx &c = ctx();
    if (!c.have_gll || !c.have_D) {
        std::vector<double> z, w, D;
        gll_build(c.nx, z, w, D);
        if (!c.have_gll) c.z_host = z, c.w_host = w, c.have_gll = true;
        if (!c.have_D) {
            c.D_host = D;
            NEKB_CUDA(cudaMemcpyToSymbolAsync(c_D, D.data(), sizeof(double) * D.size(), 0, cudaMemcpyHostToDevice, c.stream));
            c.have_D = true;
        }
    }
}

// ------------------------------------------------------------------------------------------------ numbering
struct Tuple {
    int64_t k[3];
    int64_t src;
};
inline bool tuple_less(const Tuple &a, const Tuple &b)
{
    if (a.k[0] != b.k[0]) return a.k[0] < b.k[0];
    if (a.k[1] != b.k[1]) return a.k[1] < b.k[1];
    return a.k[2] < b.k[2];
}
inline bool tuple_same(const Tuple &a, const Tuple &b) { return a.k[0] == b.k[0] && a.k[1] == b.k[1] && a.k[2] == b.k[2]; }

// core/navier8.f:1934-2002 gbtuple_rank8: a tuple lives on processor mod(key1,np) (:1964); there the distinct
// tuples are ranked in lexicographic order and offset by the number of distinct tuples on lower processors
// (:1986-1991).  rank_out[i] is the 1-based rank of t[i]; returns the global number of distinct tuples.
// With `distributed` the tuples really travel (np == nranks); otherwise the np processors are emulated.
inline int64_t rank_tuples(std::vector<Tuple> &t, int np, bool distributed, std::vector<int64_t> &rank_out)
{
    Ctx &c = ctx();
    const size_t n = t.size();
    rank_out.assign(n, 0);
    for (size_t i = 0; i < n; i++) t[i].src = (int64_t)i;
    if (!distributed) {
        std::vector<Tuple> s(t);
        std::sort(s.begin(), s.end(), [np](const Tuple &a, const Tuple &b) {
            const int64_t pa = a.k[0] % np, pb = b.k[0] % np;
            return pa != pb ? pa < pb : tuple_less(a, b);
        });
        int64_t r = 0;
        for (size_t i = 0; i < n; i++) {
            if (i == 0 || !tuple_same(s[i - 1], s[i])) r++;
            rank_out[(size_t)s[i].src] = r;
        }
        return r;
    }
    std::vector<std::vector<Tuple>> out(np);
    for (size_t i = 0; i < n; i++) out[(int)(t[i].k[0] % np)].push_back(t[i]);
    std::vector<std::vector<Tuple>> in = exchange_records(out);
    struct Held {
        Tuple t;
        int from;
        int64_t pos;
    };
    std::vector<Held> held;
    for (int r = 0; r < np; r++)
        for (size_t q = 0; q < in[r].size(); q++) held.push_back({in[r][q], r, (int64_t)q});
    std::sort(held.begin(), held.end(), [](const Held &a, const Held &b) { return tuple_less(a.t, b.t); });
    std::vector<int64_t> lrank(held.size());
    int64_t nu = 0;
    for (size_t i = 0; i < held.size(); i++) {
        if (i == 0 || !tuple_same(held[i - 1].t, held[i].t)) nu++;
        lrank[i] = nu;
    }
    std::vector<int64_t> all_nu(np);
    host_allgather(&nu, all_nu.data(), sizeof(int64_t));
    int64_t prior = 0, total = 0;
    for (int r = 0; r < np; r++) {
        if (r < c.rank) prior += all_nu[r];
        total += all_nu[r];
    }
    std::vector<std::vector<int64_t>> back(np);
    for (int r = 0; r < np; r++) back[r].resize(in[r].size());
    for (size_t i = 0; i < held.size(); i++) back[held[i].from][(size_t)held[i].pos] = lrank[i] + prior;
    std::vector<std::vector<int64_t>> got = exchange_records(back);
    // got[r][q] answers the q-th tuple this rank sent to r, i.e. out[r][q]
    for (int r = 0; r < np; r++)
        for (size_t q = 0; q < out[r].size(); q++) rank_out[(size_t)out[r][q].src] = got[r][q];
    return total;
}

// Index of the smallest of four values exactly as the reference finds it: i8rank (core/navier8.f:1131-1180,
// the Numerical Recipes `indexx` heap sort) followed by ind(1); ties are resolved the way the heap sort does.
inline int argmin4_heapsort(const int64_t a[4])
{
    int ind[4] = {1, 2, 3, 4};
    int n = 4, l = n / 2 + 1, ir = n, indx, i, j;
    int64_t q;
    for (;;) {
        if (l > 1) {
            indx = ind[--l - 1];
            q = a[indx - 1];
        } else {
            indx = ind[ir - 1];
            q = a[indx - 1];
            ind[ir - 1] = ind[0];
            if (--ir == 1) {
                ind[0] = indx;
                break;
            }
        }
        i = l;
        j = l + l;
        while (j <= ir) {
            if (j < ir && a[ind[j - 1] - 1] < a[ind[j] - 1]) j++;
            if (q < a[ind[j - 1] - 1]) {
                ind[i - 1] = ind[j - 1];
                i = j;
                j += j;
            } else
                j = ir + 1;
        }
        ind[i - 1] = indx;
    }
    return ind[0] - 1;
}

// core/navier8.f:2004-2360 setvert3d, ifcenter=.false.: vertices keep their ids, the nx-2 interior nodes of a
// unique edge get consecutive ids oriented from its smaller to its larger end vertex, the (nx-2)^2 interior
// nodes of a unique face get consecutive ids in a traversal that starts at the face's smallest vertex and runs
// first towards the smaller of its two neighbours; element interiors get 0.
inline int64_t setvert3d_host(int64_t *glo_num, int nx, int64_t nel, const int64_t *vertex, int np)
{
    Ctx &c = ctx();
    const bool distributed = c.nranks > 1 && np == c.nranks && (c.allgather != nullptr || c.nccl_comm != nullptr);
    NEKB_REQUIRE(np >= 1, "setvert3d: np must be >= 1");
    const int64_t nxyz = (int64_t)nx * nx * nx;
    const int L = nx - 1;
    auto node = [nx](int i, int j, int k) { return (int64_t)i + (int64_t)nx * (j + (int64_t)nx * k); };
    std::fill(glo_num, glo_num + nxyz * nel, (int64_t)0);
    int64_t vmax = 0;
    for (int64_t q = 0; q < 8 * nel; q++) vmax = std::max(vmax, vertex[q]);
    if (distributed) {
        std::vector<int64_t> all(c.nranks);
        host_allgather(&vmax, all.data(), sizeof(int64_t));
        for (int64_t v : all) vmax = std::max(vmax, v);
    }
    for (int64_t e = 0; e < nel; e++)
        for (int cnr = 0; cnr < 8; cnr++)
            glo_num[nxyz * e + node(L * (cnr & 1), L * ((cnr >> 1) & 1), L * (cnr >> 2))] = vertex[8 * e + cnr];
    int64_t ngv = vmax;
    if (nx == 2) return ngv;

    // ---- edges: tuple index b + 2c + 4d, d = direction of the edge, (b,c) = the two fixed coordinates -------
    const int nin = nx - 2;
    auto edge_ends = [](int d, int b, int cc, int &c0, int &c1) {
        if (d == 0) c0 = 2 * b + 4 * cc, c1 = c0 + 1;  // along r: j=b, k=c
        if (d == 1) c0 = b + 4 * cc, c1 = c0 + 2;      // along s: i=b, k=c
        if (d == 2) c0 = b + 2 * cc, c1 = c0 + 4;      // along t: i=b, j=c
    };
    {
        std::vector<Tuple> et((size_t)12 * nel);
        for (int64_t e = 0; e < nel; e++)
            for (int d = 0; d < 3; d++)
                for (int cc = 0; cc < 2; cc++)
                    for (int b = 0; b < 2; b++) {
                        int c0, c1;
                        edge_ends(d, b, cc, c0, c1);
                        const int64_t v0 = vertex[8 * e + c0], v1 = vertex[8 * e + c1];
                        Tuple &t = et[(size_t)(12 * e + b + 2 * cc + 4 * d)];
                        t.k[0] = std::min(v0, v1), t.k[1] = std::max(v0, v1), t.k[2] = 0;
                    }
        std::vector<int64_t> erank;
        const int64_t n_edges = rank_tuples(et, np, distributed, erank);
        for (int64_t e = 0; e < nel; e++)
            for (int d = 0; d < 3; d++)
                for (int cc = 0; cc < 2; cc++)
                    for (int b = 0; b < 2; b++) {
                        int c0, c1;
                        edge_ends(d, b, cc, c0, c1);
                        const bool fwd = vertex[8 * e + c0] < vertex[8 * e + c1];
                        const int64_t base = ngv + (int64_t)nin * (erank[(size_t)(12 * e + b + 2 * cc + 4 * d)] - 1);
                        for (int t = 1; t < L; t++) {
                            const int i = d == 0 ? t : L * b;
                            const int j = d == 1 ? t : (d == 0 ? L * b : L * cc);
                            const int k = d == 2 ? t : L * cc;
                            glo_num[nxyz * e + node(i, j, k)] = base + (fwd ? t : L - t);
                        }
                    }
        ngv += n_edges * nin;
    }

    // ---- faces: 1,2 = r-,r+ ; 3,4 = s-,s+ ; 5,6 = t-,t+ (core/TOPOL:37-41 icface) ------------------------------
    {
        std::vector<Tuple> ft((size_t)6 * nel);
        auto face_corner = [](int f, int a, int b) {  // corner of face f at face-local (a,b), symmetric numbering
            const int side = f & 1, dir = f >> 1;
            if (dir == 0) return side + 2 * a + 4 * b;  // local axes (s,t)
            if (dir == 1) return a + 2 * side + 4 * b;  // (r,t)
            return a + 2 * b + 4 * side;                // (r,s)
        };
        for (int64_t e = 0; e < nel; e++)
            for (int f = 0; f < 6; f++) {
                int64_t v[4];
                for (int q = 0; q < 4; q++) v[q] = vertex[8 * e + face_corner(f, q & 1, q >> 1)];
                std::sort(v, v + 4);
                Tuple &t = ft[(size_t)(6 * e + f)];
                t.k[0] = v[0], t.k[1] = v[1], t.k[2] = v[2];
            }
        std::vector<int64_t> frank;
        const int64_t n_faces = rank_tuples(ft, np, distributed, frank);
        const int64_t nin2 = (int64_t)nin * nin;
        for (int64_t e = 0; e < nel; e++)
            for (int f = 0; f < 6; f++) {
                const int side = f & 1, dir = f >> 1;
                int64_t gv[4];  // (lo,lo) (hi,lo) (lo,hi) (hi,hi) in face-local (a,b)
                for (int q = 0; q < 4; q++) gv[q] = vertex[8 * e + face_corner(f, q & 1, q >> 1)];
                const int m = argmin4_heapsort(gv);
                const bool flip_a = m & 1, flip_b = m >> 1;
                // neighbours of the smallest corner along a and along b
                const bool a_fast = gv[m ^ 1] < gv[m ^ 2];
                const int64_t base = ngv + nin2 * (frank[(size_t)(6 * e + f)] - 1);
                for (int b = 1; b < L; b++)
                    for (int a = 1; a < L; a++) {
                        const int ao = flip_a ? L - a : a, bo = flip_b ? L - b : b;
                        const int64_t l = a_fast ? (ao - 1) + (int64_t)nin * (bo - 1) : (bo - 1) + (int64_t)nin * (ao - 1);
                        int i, j, k;
                        if (dir == 0) i = L * side, j = a, k = b;
                        else if (dir == 1) i = a, j = L * side, k = b;
                        else i = a, j = b, k = L * side;
                        glo_num[nxyz * e + node(i, j, k)] = base + l + 1;
                    }
            }
        ngv += n_faces * nin2;
    }
    return ngv;
}

// ------------------------------------------------------------------------------------------------ geometry
// Arithmetic below is written with explicitly rounded operations in the order of the Fortran statements
// (no FMA contraction), so that on identical inputs the factors agree with an un-fused CPU build bit for bit.
__device__ __forceinline__ double mul_(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double add_(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double sub_(double a, double b) { return __dsub_rn(a, b); }
__device__ __forceinline__ double div_(double a, double b) { return __ddiv_rn(a, b); }

struct BoxDesc {
    int nelx, nely, nelz;  // global element counts
    int lx, ly, lz;        // this rank's brick
    int ox, oy, oz;        // its offset
    double deform;
};

// c_z / c_w (GLL points and weights in constant memory) are defined in ctx.cuh

// Box [0,1]^3, elements x-fastest (tools/genbox; examples/bp5/genbox.in:16-20); nodes by the trilinear map of
// core/genxyz.f:1269-1332 (tensr3 with the 2-point Lagrange weights (1-z)/2, (1+z)/2, contracted r, then s, then t).
template <int NX>
__global__ void __launch_bounds__(NX *NX *NX)
    box_xyz_kernel(double *__restrict__ xm1, double *__restrict__ ym1, double *__restrict__ zm1, BoxDesc b)
{
    constexpr int N3 = NX * NX * NX;
    const int e = blockIdx.x, q = threadIdx.x;
    const int i = q % NX, j = (q / NX) % NX, k = q / (NX * NX);
    const int ex = b.ox + e % b.lx, ey = b.oy + (e / b.lx) % b.ly, ez = b.oz + e / (b.lx * b.ly);
    auto corner = [](int ie, int nel) { return div_(mul_(1.0, (double)ie), (double)nel); };  // 0 + (1-0)*ie/nel
    auto blend = [](double c0, double c1, int a) {
        const double w0 = div_(sub_(1.0, c_z[a]), 2.0), w1 = div_(add_(1.0, c_z[a]), 2.0);
        return add_(mul_(w0, c0), mul_(w1, c1));
    };
    // a coordinate of a box element varies along one direction only, but the three contractions are still
    // performed so that the rounding matches the general trilinear map
    auto tri = [&](double c000, double c100, double c010, double c110, double c001, double c101, double c011,
                   double c111) {
        const double v00 = blend(c000, c100, i), v10 = blend(c010, c110, i), v01 = blend(c001, c101, i),
                     v11 = blend(c011, c111, i);
        const double w0 = blend(v00, v10, j), w1 = blend(v01, v11, j);
        return blend(w0, w1, k);
    };
    const double x0 = corner(ex, b.nelx), x1 = corner(ex + 1, b.nelx);
    const double y0 = corner(ey, b.nely), y1 = corner(ey + 1, b.nely);
    const double z0 = corner(ez, b.nelz), z1 = corner(ez + 1, b.nelz);
    double x = tri(x0, x1, x0, x1, x0, x1, x0, x1);
    double y = tri(y0, y0, y1, y1, y0, y0, y1, y1);
    double z = tri(z0, z0, z0, z0, z1, z1, z1, z1);
    if (b.deform != 0.0) {  // smooth test deformation (all six factors become non-trivial)
        const double pi = 3.14159265358979323846;
        const double s = sin(pi * x) * sin(pi * y) * sin(pi * z);
        x = x + b.deform * s;
        y = y + 0.7 * b.deform * s;
        z = z - 0.5 * b.deform * s;
    }
    const size_t o = (size_t)e * N3 + q;
    xm1[o] = x, ym1[o] = y, zm1[o] = z;
}

// Local derivatives of the three coordinates (loc_grad3 / xyzrst, three mxm's with the k = 1..n
// left-to-right sum of core/mxm_std.f), then either
//   MODE 0: examples/bp5/bp5.usr:623-699 geodatstd -> g[e][c][q], c = rr,rs,rt,ss,st,tt
//   MODE 1: core/coef.f:555-631 glmapm1 + :633-784 geodat1 -> same storage (core g1..g6 = rr,ss,tt,rs,rt,st are
//           written to slots 0,3,5,1,2,4), bm1 and jacm1.
template <int NX, int MODE>
__global__ void __launch_bounds__(NX *NX *NX)
    geom_kernel(double *__restrict__ g, double *__restrict__ bm1, double *__restrict__ jacm1,
                const double *__restrict__ xm1, const double *__restrict__ ym1, const double *__restrict__ zm1)
{
    constexpr int N2 = NX * NX, N3 = NX * NX * NX;
    __shared__ double s_x[3][N3];
    const int e = blockIdx.x, q = threadIdx.x;
    const int i = q % NX, j = (q / NX) % NX, k = q / N2;
    const size_t o = (size_t)e * N3 + q;
    s_x[0][q] = xm1[o];
    s_x[1][q] = ym1[o];
    s_x[2][q] = zm1[o];
    __syncthreads();
    double dr[3], ds[3], dt[3];
#pragma unroll
    for (int c = 0; c < 3; c++) {
        double a = 0.0, b = 0.0, d = 0.0;
        for (int m = 0; m < NX; m++) {
            a = add_(a, mul_(c_D[i * NX + m], s_x[c][k * N2 + j * NX + m]));
            b = add_(b, mul_(s_x[c][k * N2 + m * NX + i], c_D[j * NX + m]));
            d = add_(d, mul_(s_x[c][m * N2 + j * NX + i], c_D[k * NX + m]));
        }
        dr[c] = a, ds[c] = b, dt[c] = d;
    }
    const double xr = dr[0], xs = ds[0], xt = dt[0], yr = dr[1], ys = ds[1], yt = dt[1], zr = dr[2], zs = ds[2],
                 zt = dt[2];
    const double w3 = mul_(mul_(c_w[i], c_w[j]), c_w[k]);  // core/coef.f:263-267 w3m1
    double *ge = g + (size_t)e * 6 * N3 + q;
    if (MODE == 0) {
        const double jac = add_(sub_(mul_(xr, sub_(mul_(ys, zt), mul_(yt, zs))), mul_(xs, sub_(mul_(yr, zt), mul_(yt, zr)))),
                                mul_(xt, sub_(mul_(yr, zs), mul_(ys, zr))));
        const double g11 = div_(sub_(mul_(ys, zt), mul_(yt, zs)), jac), g12 = div_(sub_(mul_(xt, zs), mul_(zt, xs)), jac),
                     g13 = div_(sub_(mul_(xs, yt), mul_(ys, xt)), jac), g21 = div_(sub_(mul_(yt, zr), mul_(yr, zt)), jac),
                     g22 = div_(sub_(mul_(xr, zt), mul_(zr, xt)), jac), g23 = div_(sub_(mul_(xt, yr), mul_(yt, xr)), jac),
                     g31 = div_(sub_(mul_(yr, zs), mul_(ys, zr)), jac), g32 = div_(sub_(mul_(xs, zr), mul_(zs, xr)), jac),
                     g33 = div_(sub_(mul_(xr, ys), mul_(yr, xs)), jac);
        const double sc = mul_(w3, jac);
        auto dot3 = [](double a1, double a2, double a3, double b1, double b2, double b3) {
            return add_(add_(mul_(a1, b1), mul_(a2, b2)), mul_(a3, b3));
        };
        ge[0 * N3] = mul_(sc, dot3(g11, g12, g13, g11, g12, g13));
        ge[1 * N3] = mul_(sc, dot3(g11, g12, g13, g21, g22, g23));
        ge[2 * N3] = mul_(sc, dot3(g11, g12, g13, g31, g32, g33));
        ge[3 * N3] = mul_(sc, dot3(g21, g22, g23, g21, g22, g23));
        ge[4 * N3] = mul_(sc, dot3(g21, g22, g23, g31, g32, g33));
        ge[5 * N3] = mul_(sc, dot3(g31, g32, g33, g31, g32, g33));
        if (bm1 != nullptr) bm1[o] = sc;
    } else {
        double jac = 0.0;
        jac = add_(jac, mul_(mul_(xr, ys), zt));
        jac = add_(jac, mul_(mul_(xt, yr), zs));
        jac = add_(jac, mul_(mul_(xs, yt), zr));
        jac = sub_(jac, mul_(mul_(xr, yt), zs));
        jac = sub_(jac, mul_(mul_(xs, yr), zt));
        jac = sub_(jac, mul_(mul_(xt, ys), zr));
        const double rx = sub_(mul_(ys, zt), mul_(yt, zs)), ry = sub_(mul_(xt, zs), mul_(xs, zt)),
                     rz = sub_(mul_(xs, yt), mul_(xt, ys)), sx = sub_(mul_(yt, zr), mul_(yr, zt)),
                     sy = sub_(mul_(xr, zt), mul_(xt, zr)), sz = sub_(mul_(xt, yr), mul_(xr, yt)),
                     tx = sub_(mul_(yr, zs), mul_(ys, zr)), ty = sub_(mul_(xs, zr), mul_(xr, zs)),
                     tz = sub_(mul_(xr, ys), mul_(xs, yr));
        const double wj = div_(1.0, jac);
        auto fac = [&](double a1, double a2, double a3, double b1, double b2, double b3) {
            return mul_(mul_(add_(add_(mul_(a1, b1), mul_(a2, b2)), mul_(a3, b3)), wj), w3);
        };
        ge[0 * N3] = fac(rx, ry, rz, rx, ry, rz);  // g1 rr
        ge[3 * N3] = fac(sx, sy, sz, sx, sy, sz);  // g2 ss
        ge[5 * N3] = fac(tx, ty, tz, tx, ty, tz);  // g3 tt
        ge[1 * N3] = fac(rx, ry, rz, sx, sy, sz);  // g4 rs
        ge[2 * N3] = fac(rx, ry, rz, tx, ty, tz);  // g5 rt
        ge[4 * N3] = fac(sx, sy, sz, tx, ty, tz);  // g6 st
        if (bm1 != nullptr) bm1[o] = mul_(jac, w3);
        if (jacm1 != nullptr) jacm1[o] = jac;
    }
}

inline void upload_gll_constants()
{
    Ctx &c = ctx();
    ensure_operators();
    NEKB_CUDA(cudaMemcpyToSymbolAsync(c_z, c.z_host.data(), sizeof(double) * c.nx, 0, cudaMemcpyHostToDevice, c.stream));
    NEKB_CUDA(cudaMemcpyToSymbolAsync(c_w, c.w_host.data(), sizeof(double) * c.nx, 0, cudaMemcpyHostToDevice, c.stream));
}

// mode 0 = BP5 geodatstd, 1 = core glmapm1/geodat1.  Fills ctx().g (and bm1; jac_out optional, device).
inline void geom_from_xyz(const double *x, const double *y, const double *z, int nel, int mode, double *jac_out)
{
    Ctx &c = ctx();
    upload_gll_constants();
    c.g.alloc((size_t)6 * c.nxyz * nel);
    c.bm1.alloc((size_t)c.nxyz * nel);
    if (nel > 0) {
        switch (c.nx) {
#define NEKB_CASE(NXV)                                                                                              \
    case NXV:                                                                                                       \
        if (mode == 0)                                                                                              \
            geom_kernel<NXV, 0><<<nel, NXV * NXV * NXV, 0, c.stream>>>(c.g.p, c.bm1.p, jac_out, x, y, z);          \
        else                                                                                                        \
            geom_kernel<NXV, 1><<<nel, NXV * NXV * NXV, 0, c.stream>>>(c.g.p, c.bm1.p, jac_out, x, y, z);          \
        break;
            NEKB_CASE(2) NEKB_CASE(3) NEKB_CASE(4) NEKB_CASE(5) NEKB_CASE(6) NEKB_CASE(7) NEKB_CASE(8) NEKB_CASE(10)
#undef NEKB_CASE
            default: NEKB_REQUIRE(false, "unsupported lx1 for geometry (supported: 2-8, 10)");
        }
        NEKB_LAUNCHED();
    }
    c.have_geom = true;
    c.geom_gen++;
    c.nelt = nel;
    if (c.nelv == 0 || c.nelv > nel) c.nelv = nel;
}

// ------------------------------------------------------------------------------------------------ BP5 case
// core/navier5.f:2650-2684 ran1: Park-Miller minimal standard generator with Bays-Durham shuffle
// (Numerical Recipes 2nd ed., p. 271), seeded as the first call of a fresh process does: idum = max(-idum,1)
// with iy = 0, i.e. seed 1 whatever the argument (navier5.f:2665-2666).
inline void ran1_stream(double *x, int64_t n)
{
    const int IA = 16807, IM = 2147483647, IQ = 127773, IR = 2836, NTAB = 32, NDIV = 1 + (IM - 1) / NTAB;
    const double AM = 1.0 / IM, RNMX = 1.0 - 1.2e-7;
    int iv[NTAB], iy, idum = 1;
    for (int j = NTAB + 7; j >= 0; j--) {
        const int k = idum / IQ;
        idum = IA * (idum - k * IQ) - IR * k;
        if (idum < 0) idum += IM;
        if (j < NTAB) iv[j] = idum;
    }
    iy = iv[0];
    for (int64_t t = 0; t < n; t++) {
        const int k = idum / IQ;
        idum = IA * (idum - k * IQ) - IR * k;
        if (idum < 0) idum += IM;
        const int j = iy / NDIV;
        iy = iv[j];
        iv[j] = idum;
        const double r = AM * iy;
        x[t] = r < RNMX ? r : RNMX;
    }
}

// bp5.usr:142-153 xmask1 target: homogeneous Dirichlet on all six sides of the box ('v  ' in genbox.in).
template <int NX>
__global__ void __launch_bounds__(NX *NX *NX) box_mask_kernel(double *__restrict__ mask, BoxDesc b)
{
    constexpr int N3 = NX * NX * NX, L = NX - 1;
    const int e = blockIdx.x, q = threadIdx.x;
    const int i = q % NX, j = (q / NX) % NX, k = q / (NX * NX);
    const int ex = b.ox + e % b.lx, ey = b.oy + (e / b.lx) % b.ly, ez = b.oz + e / (b.lx * b.ly);
    const bool bnd = (ex == 0 && i == 0) || (ex == b.nelx - 1 && i == L) || (ey == 0 && j == 0) ||
                     (ey == b.nely - 1 && j == L) || (ez == 0 && k == 0) || (ez == b.nelz - 1 && k == L);
    mask[(size_t)e * N3 + q] = bnd ? 0.0 : 1.0;
}

__global__ void __launch_bounds__(256) fill_kernel(double *__restrict__ a, double v, int64_t n)
{
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < n; t += (int64_t)gridDim.x * blockDim.x) a[t] = v;
}

struct Bp5Case {
    bool built = false;
    BoxDesc box{};
    int64_t nel = 0, n = 0;
    int gs_handle = -1;
    DevBuf<double> xm1, ym1, zm1, e1, r1, u1, mask, mult;
    DevBuf<int64_t> glo_num;
    std::vector<int64_t> vertex;
    int64_t ngv = 0;
};
inline Bp5Case &bp5case()
{
    static Bp5Case b;
    return b;
}

int gs_setup_from_host_ids(const int64_t *id_host, int64_t n, const int32_t *cand, int64_t ncand);  // nekb200.cu

inline void bp5_setup(int nelx, int nely, int nelz, int px, int py, int pz, double deform)
{
    Ctx &c = ctx();
    Bp5Case &b = bp5case();
    cudaStream_t s = c.stream;
    NEKB_REQUIRE(px * py * pz == c.nranks, "bp5_setup: px*py*pz must equal the number of ranks");
    NEKB_REQUIRE(nelx % px == 0 && nely % py == 0 && nelz % pz == 0, "bp5_setup: bricks must divide the box");
    NEKB_REQUIRE(c.nx >= 2, "bp5_setup: lx1 must be >= 2");
    if (b.gs_handle >= 0) {
        gs_release(gs_get(b.gs_handle));
        b.gs_handle = -1;
    }
    BoxDesc d;
    d.nelx = nelx, d.nely = nely, d.nelz = nelz;
    d.lx = nelx / px, d.ly = nely / py, d.lz = nelz / pz;
    const int rx = c.rank % px, ry = (c.rank / px) % py, rz = c.rank / (px * py);
    d.ox = rx * d.lx, d.oy = ry * d.ly, d.oz = rz * d.lz;
    d.deform = deform;
    b.box = d;
    b.nel = (int64_t)d.lx * d.ly * d.lz;
    b.n = b.nel * c.nxyz;
    NEKB_REQUIRE(b.n < (int64_t)2147483647, "bp5_setup: too many local nodes for int32 indexing");
    const int nel = (int)b.nel;
    const int64_t n = b.n;
    const int nx = c.nx, L = nx - 1;
    upload_gll_constants();

    // coordinates, geometric factors, mask -------------------------------------------------------------
    b.xm1.alloc(n), b.ym1.alloc(n), b.zm1.alloc(n), b.mask.alloc(n);
    switch (nx) {
#define NEKB_CASE(NXV)                                                                                     \
    case NXV:                                                                                              \
        box_xyz_kernel<NXV><<<nel, NXV * NXV * NXV, 0, s>>>(b.xm1.p, b.ym1.p, b.zm1.p, d);                \
        box_mask_kernel<NXV><<<nel, NXV * NXV * NXV, 0, s>>>(b.mask.p, d);                                \
        break;
        NEKB_CASE(2) NEKB_CASE(3) NEKB_CASE(4) NEKB_CASE(5) NEKB_CASE(6) NEKB_CASE(7) NEKB_CASE(8) NEKB_CASE(10)
#undef NEKB_CASE
        default: NEKB_REQUIRE(false, "unsupported lx1 for the BP5 case (supported: 2-8, 10)");
    }
    launch_counter() += 2;
    NEKB_CUDA(cudaPeekAtLastError());
    geom_from_xyz(b.xm1.p, b.ym1.p, b.zm1.p, nel, 0, nullptr);
    c.nelv = c.nelt = nel;
    c.v1mask.alloc(n);
    NEKB_CUDA(cudaMemcpyAsync(c.v1mask.p, b.mask.p, sizeof(double) * n, cudaMemcpyDeviceToDevice, s));

    // numbering (host) ------------------------------------------------------------------------------------------
    b.vertex.resize((size_t)8 * nel);
    const int64_t nvx = nelx + 1, nvy = nely + 1;
    for (int e = 0; e < nel; e++) {
        const int ex = d.ox + e % d.lx, ey = d.oy + (e / d.lx) % d.ly, ez = d.oz + e / (d.lx * d.ly);
        for (int q = 0; q < 8; q++)
            b.vertex[(size_t)8 * e + q] = 1 + (ex + (q & 1)) + nvx * ((ey + ((q >> 1) & 1)) + nvy * (int64_t)(ez + (q >> 2)));
    }
    std::vector<int64_t> glo((size_t)n);
    b.ngv = setvert3d_host(glo.data(), nx, nel, b.vertex.data(), c.nranks);
    // BASELINE config 3 (E swept until HBM is full): above a million elements (or with NEKB_BP5_LEAN=1) everything the solve
    // itself does not read is released or never uploaded -- coordinates, the int64 numbering, the mass matrix: 5 of 15
    // field-sized arrays.  nekb_bp5_get / nekb_bp5_devptr on a released array fail loudly.
    const char *lean_env = getenv("NEKB_BP5_LEAN");
    const bool lean = lean_env ? atoi(lean_env) != 0 : nel > 1000000;
    if (!lean) b.glo_num.upload(glo.data(), (size_t)n, s);
    else b.xm1.release(), b.ym1.release(), b.zm1.release();      // before the gs set-up's temporaries are allocated
    // candidates for sharing with other ranks: nodes on brick faces that have a neighbouring brick
    std::vector<int32_t> cand;
    if (c.nranks > 1) {
        const bool fx0 = rx > 0, fx1 = rx < px - 1, fy0 = ry > 0, fy1 = ry < py - 1, fz0 = rz > 0, fz1 = rz < pz - 1;
        for (int e = 0; e < nel; e++) {
            const int ex = e % d.lx, ey = (e / d.lx) % d.ly, ez = e / (d.lx * d.ly);
            const bool ax0 = fx0 && ex == 0, ax1 = fx1 && ex == d.lx - 1, ay0 = fy0 && ey == 0,
                       ay1 = fy1 && ey == d.ly - 1, az0 = fz0 && ez == 0, az1 = fz1 && ez == d.lz - 1;
            if (!(ax0 || ax1 || ay0 || ay1 || az0 || az1)) continue;
            for (int k = 0; k < nx; k++)
                for (int j = 0; j < nx; j++)
                    for (int i = 0; i < nx; i++)
                        if ((ax0 && i == 0) || (ax1 && i == L) || (ay0 && j == 0) || (ay1 && j == L) || (az0 && k == 0) ||
                            (az1 && k == L))
                            cand.push_back((int32_t)((int64_t)e * c.nxyz + i + nx * (j + nx * k)));
        }
    }
    b.gs_handle = gs_setup_from_host_ids(glo.data(), n, c.nranks > 1 ? cand.data() : nullptr, (int64_t)cand.size());
    c.gsh_fld[c.ifield] = b.gs_handle;

    // multiplicity: vmult = 1/dssum(1) (core/connect1.f:124-135) ----------------------------------------------
    b.mult.alloc(n);
    fill_kernel<<<blocks_for(n), 256, 0, s>>>(b.mult.p, 1.0, n);
    NEKB_LAUNCHED();
    gs_op(b.gs_handle, b.mult.p, 1, nullptr);
    invcol1_kernel<<<cg_grid(n), CG_THREADS, 0, s>>>(b.mult.p, n);
    NEKB_LAUNCHED();

    // exact solution e1 = mask * dsavg(ran1 field); rhs r1 = mask * dssum(A e1) (bp5.usr:352-360) ---------------
    {
        std::vector<double> rnd((size_t)n);
        ran1_stream(rnd.data(), n);
        b.e1.upload(rnd.data(), (size_t)n, s);
        NEKB_CUDA(cudaStreamSynchronize(s));
    }
    gs_op(b.gs_handle, b.e1.p, 1, nullptr);
    col2_kernel<<<blocks_for(n), 256, 0, s>>>(b.e1.p, b.mult.p, n);  // dsavg: dssum then * vmult
    col2_kernel<<<blocks_for(n), 256, 0, s>>>(b.e1.p, b.mask.p, n);
    launch_counter() += 2;
    b.r1.alloc(n);
    launch_ax(b.e1.p, b.r1.p, nullptr, nullptr, nel, nullptr);
    gs_op(b.gs_handle, b.r1.p, 1, b.mask.p);
    b.u1.alloc(n);
    b.u1.zero(s);
    NEKB_CUDA(cudaStreamSynchronize(s));
    if (lean) c.bm1.release();
    b.built = true;
}

}  // namespace nekb
