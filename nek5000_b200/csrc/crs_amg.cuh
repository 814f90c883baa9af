// crs_amg.cuh -- HOST set-up of an aggregation hierarchy for coarse (vertex-mesh) problems that are too large for the dense
// inverse of hsmg.cuh (> NEKB_CRS_DENSE_MAX distinct vertices).
//
// Role in the reference: crs_solve (core/fcrs.c:80 -> core/crs_xxt.c:926-965, or core/crs_amg.c with param(40) = 1) solves
// the assembled vertex-mesh system directly.  Plan here (prototype scripts/proto_coarse_amg.py): CG to
// rounding level preconditioned by one cycle over this hierarchy -- greedy aggregation on the strength graph, piecewise-
// constant prolongation (optionally smoothed by one damped-Jacobi step: smoothed aggregation), Galerkin coarse matrices,
// damped Jacobi, the dense inverse at the coarsest level.
//
// This file is the set-up and runs on the host (like setvert3d_host and gen_fast); it is exercised on the CPU by
// tests/test_crs_amg_host.py through nekb_crs_amg_*.  The device cycle that consumes it is crs_amg_dev.cuh.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <numeric>
#include <vector>

#include "common.cuh"

namespace nekb {

struct CsrHost {
    int64_t n = 0;
    int64_t ncols = 0;             // 0 = square
    std::vector<int64_t> rowptr;   // [n+1]
    std::vector<int32_t> col;      // ascending inside a row
    std::vector<double> val;
    int64_t nnz() const { return (int64_t)col.size(); }
};

// Sums duplicate (i,j) entries in INPUT order (stable sort), drops nothing: explicit zeros of the element matrices stay in
// the pattern, exactly as an assembled element-by-element operator has them.
inline CsrHost csr_from_triplets(int64_t n, const std::vector<int64_t> &I, const std::vector<int64_t> &J, const std::vector<double> &V)
{
    NEKB_REQUIRE(I.size() == J.size() && I.size() == V.size(), "csr_from_triplets: ragged input");
    std::vector<int64_t> ord(I.size());
    std::iota(ord.begin(), ord.end(), (int64_t)0);
    std::stable_sort(ord.begin(), ord.end(), [&](int64_t a, int64_t b) { return I[a] != I[b] ? I[a] < I[b] : J[a] < J[b]; });
    CsrHost A;
    A.n = n;
    A.rowptr.assign((size_t)n + 1, 0);
    int64_t pi = -1, pj = -1;
    for (int64_t q : ord) {
        NEKB_REQUIRE(I[q] >= 0 && I[q] < n && J[q] >= 0 && J[q] < n, "csr_from_triplets: index outside the matrix");
        if (I[q] == pi && J[q] == pj) {
            A.val.back() += V[q];
        } else {
            A.col.push_back((int32_t)J[q]);
            A.val.push_back(V[q]);
            A.rowptr[(size_t)I[q] + 1]++;
            pi = I[q], pj = J[q];
        }
    }
    for (int64_t i = 0; i < n; i++) A.rowptr[i + 1] += A.rowptr[i];
    return A;
}

// Greedy aggregation on the strength graph |a_ij| >= theta sqrt(a_ii a_jj), i != j: in index order, a free vertex whose
// strong neighbours are all free becomes a root and takes them; the leftovers then, again in index order, join the aggregate
// of their strongest aggregated strong neighbour (first one on ties, column order; leftovers placed earlier count), or start
// their own.
inline std::vector<int32_t> amg_aggregate(const CsrHost &A, double theta, int32_t &na)
{
    const int64_t n = A.n;
    std::vector<double> d((size_t)n, 0.0);
    for (int64_t i = 0; i < n; i++)
        for (int64_t q = A.rowptr[i]; q < A.rowptr[i + 1]; q++)
            if (A.col[q] == i) d[i] = A.val[q];
    auto strong = [&](int64_t i, int64_t q) {
        const int32_t j = A.col[q];
        return j != i && std::fabs(A.val[q]) >= theta * std::sqrt(d[i] * d[j]);
    };
    std::vector<int32_t> agg((size_t)n, -1);
    na = 0;
    // Decoupled unknowns (a row with no off-diagonal non-zero: the identity rows of masked / Dirichlet vertices) would each
    // stay an aggregate of their own on every level and keep the hierarchy from ever reaching `nmax` (a box side of
    // Dirichlet vertices: 49^2 singletons at 48^3 elements).  They all share ONE aggregate: its coarse row is the sum of
    // their diagonals (still SPD), and their right-hand side is zero wherever the solver is used (masked dofs).
    int32_t iso = -1;
    for (int64_t i = 0; i < n; i++) {
        bool coupled = false;
        for (int64_t q = A.rowptr[i]; q < A.rowptr[i + 1] && !coupled; q++) coupled = A.col[q] != i && A.val[q] != 0.0;
        if (!coupled) {
            if (iso < 0) iso = na++;
            agg[i] = iso;
        }
    }
    for (int64_t i = 0; i < n; i++) {
        if (agg[i] >= 0) continue;
        bool all_free = true;
        for (int64_t q = A.rowptr[i]; q < A.rowptr[i + 1] && all_free; q++)
            if (strong(i, q) && agg[A.col[q]] >= 0) all_free = false;
        if (!all_free) continue;
        agg[i] = na;
        for (int64_t q = A.rowptr[i]; q < A.rowptr[i + 1]; q++)
            if (strong(i, q)) agg[A.col[q]] = na;
        na++;
    }
    for (int64_t i = 0; i < n; i++) {
        if (agg[i] >= 0) continue;
        double best = -1.0;
        int32_t to = -1;
        for (int64_t q = A.rowptr[i]; q < A.rowptr[i + 1]; q++)
            if (strong(i, q) && agg[A.col[q]] >= 0 && std::fabs(A.val[q]) > best) best = std::fabs(A.val[q]), to = agg[A.col[q]];
        if (to < 0) to = na++;
        agg[i] = to;
    }
    return agg;
}

// A_c = P^T A P for the piecewise-constant P of `agg` (A_c[I][J] = sum of a_ij over i in I, j in J), summed in row-major
// order of the fine matrix.
inline CsrHost amg_galerkin(const CsrHost &A, const std::vector<int32_t> &agg, int32_t na)
{
    std::vector<int64_t> I, J;
    std::vector<double> V;
    I.reserve((size_t)A.nnz()), J.reserve((size_t)A.nnz()), V.reserve((size_t)A.nnz());
    for (int64_t i = 0; i < A.n; i++)
        for (int64_t q = A.rowptr[i]; q < A.rowptr[i + 1]; q++) I.push_back(agg[i]), J.push_back(agg[A.col[q]]), V.push_back(A.val[q]);
    return csr_from_triplets(na, I, J, V);
}

// C = A B (Gustavson: one dense accumulator row, columns of a result row emitted in ascending order; products are added in
// the traversal order of A's and B's rows, so the result is reproducible).
inline CsrHost spgemm(const CsrHost &A, const CsrHost &B)
{
    const int64_t bc = B.ncols ? B.ncols : B.n;
    NEKB_REQUIRE((A.ncols ? A.ncols : A.n) == B.n, "spgemm: inner dimensions differ");
    CsrHost C;
    C.n = A.n, C.ncols = bc;
    C.rowptr.assign((size_t)A.n + 1, 0);
    std::vector<double> acc((size_t)bc, 0.0);
    std::vector<int64_t> mark((size_t)bc, -1);
    std::vector<int32_t> cols;
    for (int64_t i = 0; i < A.n; i++) {
        cols.clear();
        for (int64_t q = A.rowptr[i]; q < A.rowptr[i + 1]; q++) {
            const int32_t k = A.col[q];
            const double a = A.val[q];
            for (int64_t t = B.rowptr[k]; t < B.rowptr[k + 1]; t++) {
                const int32_t j = B.col[t];
                if (mark[j] != i) mark[j] = i, acc[j] = 0.0, cols.push_back(j);
                acc[j] += a * B.val[t];
            }
        }
        std::sort(cols.begin(), cols.end());
        for (int32_t j : cols) C.col.push_back(j), C.val.push_back(acc[j]);
        C.rowptr[i + 1] = (int64_t)C.col.size();
    }
    return C;
}

inline CsrHost csr_transpose(const CsrHost &A)
{
    const int64_t nc = A.ncols ? A.ncols : A.n;
    CsrHost T;
    T.n = nc, T.ncols = A.n;
    T.rowptr.assign((size_t)nc + 1, 0);
    for (int32_t j : A.col) T.rowptr[(size_t)j + 1]++;
    for (int64_t j = 0; j < nc; j++) T.rowptr[j + 1] += T.rowptr[j];
    T.col.resize(A.col.size()), T.val.resize(A.val.size());
    std::vector<int64_t> cur(T.rowptr.begin(), T.rowptr.end() - 1);
    for (int64_t i = 0; i < A.n; i++)
        for (int64_t q = A.rowptr[i]; q < A.rowptr[i + 1]; q++) {
            const int64_t d = cur[A.col[q]]++;
            T.col[d] = (int32_t)i, T.val[d] = A.val[q];
        }
    return T;
}

// Prolongation of a level: the tentative (piecewise-constant) one, or with omega_p > 0 its smoothed form
// P = (I - omega_p D^-1 A) P_tent (smoothed aggregation).
inline CsrHost amg_prolongator(const CsrHost &A, const std::vector<int32_t> &agg, int32_t na, double omega_p)
{
    CsrHost T;
    T.n = A.n, T.ncols = na;
    T.rowptr.resize((size_t)A.n + 1);
    T.col.assign(agg.begin(), agg.end());
    T.val.assign((size_t)A.n, 1.0);
    std::iota(T.rowptr.begin(), T.rowptr.end(), (int64_t)0);
    if (!(omega_p > 0.0)) return T;
    CsrHost AT = spgemm(A, T);
    CsrHost P;
    P.n = A.n, P.ncols = na;
    P.rowptr.assign((size_t)A.n + 1, 0);
    for (int64_t i = 0; i < A.n; i++) {
        double d = 0.0;
        for (int64_t q = A.rowptr[i]; q < A.rowptr[i + 1]; q++)
            if (A.col[q] == i) d = A.val[q];
        NEKB_REQUIRE(d != 0.0, "amg_prolongator: zero diagonal");
        bool own = false;
        for (int64_t q = AT.rowptr[i]; q < AT.rowptr[i + 1]; q++) {   // ascending columns; the tentative entry merges in
            double v = -omega_p * AT.val[q] / d;
            if (AT.col[q] == agg[i]) v += 1.0, own = true;
            P.col.push_back(AT.col[q]), P.val.push_back(v);
        }
        NEKB_REQUIRE(own, "amg_prolongator: a row of A T misses its own aggregate");   // a_ii != 0 guarantees the entry
        P.rowptr[i + 1] = (int64_t)P.col.size();
    }
    return P;
}

struct AmgHierarchy {
    std::vector<CsrHost> A;                   // A[0] = the fine operator ... A.back() = the coarsest (dense inverse)
    std::vector<std::vector<int32_t>> agg;    // agg[l][i] = aggregate (row of A[l+1]) of row i of A[l]
    std::vector<CsrHost> P, PT;               // prolongation of level l (n_l x n_{l+1}) and its transpose
};

inline AmgHierarchy amg_build(CsrHost A0, int64_t nmax, double theta, double omega_p = 0.0)
{
    NEKB_REQUIRE(nmax >= 1 && theta > 0.0 && omega_p >= 0.0, "amg_build: bad parameters");
    AmgHierarchy H;
    H.A.push_back(std::move(A0));
    while (H.A.back().n > nmax) {
        int32_t na = 0;
        std::vector<int32_t> g = amg_aggregate(H.A.back(), theta, na);
        if (!((double)na < 0.7 * (double)H.A.back().n)) {
            // no further coarsening at this strength threshold: this level becomes the coarsest one if the dense inverse that
            // serves it can hold it, otherwise the set-up fails loudly
            NEKB_REQUIRE(H.A.back().n <= 16384 && H.A.size() > 1,
                         "amg_build: aggregation stalled (no strong connections at this theta) above the dense-solver size");
            break;
        }
        CsrHost P = amg_prolongator(H.A.back(), g, na, omega_p);
        CsrHost PT = csr_transpose(P);
        CsrHost Ac = omega_p > 0.0 ? spgemm(PT, spgemm(H.A.back(), P)) : amg_galerkin(H.A.back(), g, na);
        Ac.ncols = 0;
        H.agg.push_back(std::move(g));
        H.P.push_back(std::move(P));
        H.PT.push_back(std::move(PT));
        H.A.push_back(std::move(Ac));
    }
    return H;
}

inline AmgHierarchy &amg_host_hierarchy()
{
    static AmgHierarchy h;
    return h;
}

}  // namespace nekb
