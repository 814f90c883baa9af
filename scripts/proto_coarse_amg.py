"""DESIGN PROTOTYPE (numpy / scipy, CPU; not product code, not the oracle): iteration counts of candidate solvers for the
vertex-mesh coarse problem of h1mg_solve when it is too large for the dense inverse (> 12288 vertices).

The reference solves this system directly (XXT, core/crs_xxt.c), so a replacement must reach rounding level (1e-13 relative
residual) -- the question is how many sweeps over the sparse matrix that takes.

  * jacobi   : the current device path, Jacobi-preconditioned CG;
  * agg-V    : CG preconditioned by one V(1,1) cycle of a plain-aggregation hierarchy (piecewise-constant prolongation built
               by greedy aggregation on the matrix graph, Galerkin coarse matrices, damped-Jacobi smoothing, the existing
               dense inverse at the coarsest level of <= 4096 unknowns);
  * agg-K    : the same with a two-step Krylov (K-)cycle on the second level;
  * smoothed : the same aggregates with the prolongation smoothed by one damped-Jacobi step (smoothed aggregation).

Operator: Q1 stiffness matrix on an m^3 box (the Galerkin projection of the SEM operator onto the element-corner basis is
spectrally equivalent), Neumann walls, one Dirichlet (outflow) side.      python scripts/proto_coarse_amg.py 16 32 48
"""
import sys
import time

import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla


def q1_stiffness(m):
    """Assembled Q1 Laplacian on (m+1)^3 vertices of a unit-spaced box; x = m side Dirichlet (rows/cols removed)."""
    g = np.array([-1, 1]) / np.sqrt(3.0)
    K = np.zeros((8, 8))
    for a in g:
        for b in g:
            for c in g:
                dN = np.zeros((8, 3))
                for q in range(8):
                    sx, sy, sz = (2 * (q & 1) - 1), (2 * ((q >> 1) & 1) - 1), (2 * (q >> 2) - 1)
                    dN[q] = [sx * (1 + sy * b) * (1 + sz * c) / 8, sy * (1 + sx * a) * (1 + sz * c) / 8, sz * (1 + sx * a) * (1 + sy * b) / 8]
                K += (dN @ dN.T) * 2.0          # jacobian (1/2)^3, inverse metric 2^2 -> factor 1/2 * 4 = 2
    n1 = m + 1
    e = np.arange(m ** 3)
    ex, ey, ez = e % m, (e // m) % m, e // (m * m)
    q = np.arange(8)
    vid = (ex[:, None] + (q & 1)) + n1 * ((ey[:, None] + ((q >> 1) & 1)) + n1 * (ez[:, None] + (q >> 2)))
    I = np.repeat(vid[:, :, None], 8, axis=2).ravel()
    J = np.repeat(vid[:, None, :], 8, axis=1).ravel()
    A = sp.coo_matrix((np.tile(K.ravel(), m ** 3), (I, J)), shape=(n1 ** 3, n1 ** 3)).tocsr()
    keep = np.where((np.arange(n1 ** 3) % n1) != m)[0]
    return A[keep][:, keep].tocsr()


def aggregate(A, theta=0.02):
    """Greedy aggregation on the strength graph (|a_ij| >= theta sqrt(a_ii a_jj)): roots take their free strong neighbours,
    leftovers join the neighbouring aggregate they are most strongly tied to."""
    n = A.shape[0]
    d = A.diagonal()
    C = A.tocoo()
    strong = (C.row != C.col) & (np.abs(C.data) >= theta * np.sqrt(d[C.row] * d[C.col]))
    S = sp.csr_matrix((np.abs(C.data[strong]), (C.row[strong], C.col[strong])), shape=(n, n))
    agg = -np.ones(n, dtype=np.int64)
    na = 0
    # decoupled unknowns (no off-diagonal non-zero: identity rows of masked vertices) share one aggregate
    offd = sp.csr_matrix(((C.data != 0) & (C.row != C.col), (C.row, C.col)), shape=(n, n))
    iso = np.asarray(offd.sum(axis=1)).ravel() == 0
    if iso.any():
        agg[iso] = 0
        na = 1
    ip, ix = S.indptr, S.indices
    for i in range(n):
        if agg[i] >= 0:
            continue
        nb = ix[ip[i]:ip[i + 1]]
        if np.all(agg[nb] < 0):
            agg[i] = na
            agg[nb] = na
            na += 1
    for i in np.where(agg < 0)[0]:
        nb, w = ix[ip[i]:ip[i + 1]], S.data[ip[i]:ip[i + 1]]
        ok = agg[nb] >= 0
        if ok.any():
            agg[i] = agg[nb[ok][np.argmax(w[ok])]]
        else:
            agg[i] = na
            na += 1
    return agg, na


class Hierarchy:
    def __init__(self, A, nmax=4096, omega=0.7):
        self.levels = []
        while A.shape[0] > nmax:
            agg, na = aggregate(A)
            assert na < 0.7 * A.shape[0], "aggregation stalled"
            P = sp.csr_matrix((np.ones(A.shape[0]), (np.arange(A.shape[0]), agg)), shape=(A.shape[0], na))
            self.levels.append((A, P, omega / A.diagonal()))
            A = (P.T @ A @ P).tocsr()
        self.Ainv = np.linalg.inv(A.toarray())
        self.sizes = [l[0].shape[0] for l in self.levels] + [A.shape[0]]
        self.nnz = [l[0].nnz for l in self.levels]

    def cycle(self, b, l=0, kcycle=False):
        if l == len(self.levels):
            return self.Ainv @ b
        A, P, dj = self.levels[l]
        x = dj * b
        r = P.T @ (b - A @ x)
        if kcycle and l == 0 and l + 1 < len(self.levels):        # two flexible-CG steps on level 1
            A1 = self.levels[1][0]
            c1 = self.cycle(r, l + 1)
            v1 = A1 @ c1
            a1 = (c1 @ r) / (c1 @ v1)
            r2 = r - a1 * v1
            c2 = self.cycle(r2, l + 1)
            v2 = A1 @ c2
            g = (c2 @ v1) / (c1 @ v1)
            c2, v2 = c2 - g * c1, v2 - g * v1
            a2 = (c2 @ r2) / (c2 @ v2)
            e = a1 * c1 + a2 * c2
        else:
            e = self.cycle(r, l + 1)
        x = x + P @ e
        return x + dj * (b - A @ x)


def smoothed_hierarchy_cycle(A, nmax=2048, omega_p=0.66, omega=0.7):
    """The same aggregates with a SMOOTHED prolongation P = (I - omega_p D^-1 A) P_tentative (classical smoothed aggregation):
    returns (cycle function, level sizes, operator complexity, nnz per row of the finest P)."""
    lev = []
    A0 = A
    while A.shape[0] > nmax:
        agg, na = aggregate(A)
        T = sp.csr_matrix((np.ones(A.shape[0]), (np.arange(A.shape[0]), agg)), shape=(A.shape[0], na))
        Pm = (T - omega_p * (sp.diags(1.0 / A.diagonal()) @ (A @ T))).tocsr()
        lev.append((A, Pm))
        A = (Pm.T @ A @ Pm).tocsr()
    Ainv = np.linalg.inv(A.toarray())

    def cycle(b, l=0):
        if l == len(lev):
            return Ainv @ b
        Al, Pm = lev[l]
        dj = omega / Al.diagonal()
        x = dj * b
        x = x + Pm @ cycle(Pm.T @ (b - Al @ x), l + 1)
        return x + dj * (b - Al @ x)

    sizes = [l[0].shape[0] for l in lev] + [A.shape[0]]
    return cycle, sizes, sum(l[0].nnz for l in lev) / max(A0.nnz, 1), (lev[0][1].nnz / A0.shape[0]) if lev else 0.0


def pcg(A, b, prec, tol=1e-13, maxit=3000):
    x = np.zeros_like(b)
    r = b.copy()
    z = prec(r)
    p = z.copy()
    rz = r @ z
    n0 = np.linalg.norm(b)
    for it in range(1, maxit + 1):
        w = A @ p
        a = rz / (p @ w)
        x += a * p
        r -= a * w
        if np.linalg.norm(r) <= tol * n0:
            return x, it
        z = prec(r)
        rz, rz0 = r @ z, rz
        p = z + (rz / rz0) * p
    return x, maxit


def main():
    ms = [int(v) for v in sys.argv[1:]] or [16, 32]
    print(f"{'m':>4} {'vertices':>9} | {'jacobi':>7} | {'levels':>24} {'agg-V':>6} {'agg-K':>6} | sparse work per solve (matrix nnz sweeps): jacobi, agg-V")
    for m in ms:
        A = q1_stiffness(m)
        n = A.shape[0]
        b = A @ np.random.default_rng(0).standard_normal(n)
        dinv = 1.0 / A.diagonal()
        _, itj = pcg(A, b, lambda r: dinv * r)
        t0 = time.time()
        H = Hierarchy(A)
        ts = time.time() - t0
        _, itv = pcg(A, b, lambda r: H.cycle(r))
        _, itk = pcg(A, b, lambda r: H.cycle(r, kcycle=True))
        # one V(1,1) cycle sweeps every level's matrix twice (residual + post-smoothing residual); CG adds one fine sweep
        wv = itv * (1 + 2 * sum(H.nnz) / A.nnz)
        print(f"{m:4d} {n:9d} | {itj:7d} | {str(H.sizes):>24} {itv:6d} {itk:6d} | {itj:.0f}, {wv:.0f}   (setup {ts:.1f} s)")
        cyc, sizes, cx, pn = smoothed_hierarchy_cycle(A)
        _, its = pcg(A, b, cyc)
        print(f"{'':>14} | smoothed aggregation {sizes}: {its} iterations, operator complexity {cx:.2f}, {pn:.1f} nnz per row of P")


if __name__ == "__main__":
    main()
