/*
 * oracle/nek_oracle.c -- CPU restatement of the Nek5000 BP5 / Helmholtz-PCG-dssum hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under nek5000_b200/ may include, link or
 * call this file; only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference leg use it, and only as the checker / the
 * CPU arm, never as the shipped path.
 *
 * PARITY PINNED AGAINST THE REFERENCE ITSELF: no Fortran compiler exists in the
 * authoring container and gslib v1.0.9 is not vendored, so the reference cannot
 * be built with its own toolchain; instead oracle/f77c.py transpiles the
 * reference's own Fortran for this path (read where it lies under
 * /root/reference) to C, oracle/ref_build.py compiles it (gcc -O2
 * -ffp-contract=off) with serial stand-ins for gslib / crs (oracle/ref_stubs.c)
 * into oracle/_ref/, and tests/test_ref_pins.py holds this restatement to the
 * reference's outputs (tests/golden/ref_golden.npz): BIT FOR BIT for speclib,
 * numbering, geometry, masks, axhelm, setprec, the gs ops, cggo and the whole
 * BP5 driver.  What stays a stand-in: the gather-scatter and the coarse solve
 * are single-rank restatements of gslib's / XXT's published semantics.
 * Additional pins: the reference's mesh fixtures examples/bp5/bp5.{re2,ma2}
 * (tests/golden) and the analytic known-answer tests in tests/test_oracle.py.
 *
 * Conventions: all arrays are Fortran column-major, u(i,j,k,e) lives at
 * u[i + nx*(j + nx*k) + nx^3*e] with 0-based i,j,k,e.  "real" is double
 * (makenek.inc promotes with -fdefault-real-8).  Build with
 * -O2 -ffp-contract=off so products and sums round like un-fused SSE2 code.
 *
 * Every function cites the reference file:line it restates (paths relative to
 * the Nek5000 tree).
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define NKO_MAXN 84

/* ------------------------------------------------------------------------- */
/* speclib.f : Gauss-Lobatto-Legendre points, weights, derivative matrix      */
/* ------------------------------------------------------------------------- */

/* core/speclib.f:479-518 JACOBF */
static void jacobf(double *poly, double *pder, double *polym1, double *pderm1,
                   double *polym2, double *pderm2, int n, double alp, double bet, double x)
{
    double apb = alp + bet;
    double polyl, pderl, psave = 0.0, pdsave = 0.0;
    *poly = 1.0;
    *pder = 0.0;
    if (n == 0) return;
    polyl = *poly;
    pderl = *pder;
    *poly = (alp - bet + (apb + 2.0) * x) / 2.0;
    *pder = (apb + 2.0) / 2.0;
    if (n == 1) return;
    for (int k = 2; k <= n; k++) {
        double dk = (double)k;
        double a1 = 2.0 * dk * (dk + apb) * (2.0 * dk + apb - 2.0);
        double a2 = (2.0 * dk + apb - 1.0) * (alp * alp - bet * bet);
        double b3 = (2.0 * dk + apb - 2.0);
        double a3 = b3 * (b3 + 1.0) * (b3 + 2.0);
        double a4 = 2.0 * (dk + alp - 1.0) * (dk + bet - 1.0) * (2.0 * dk + apb);
        double polyn = ((a2 + a3 * x) * (*poly) - a4 * polyl) / a1;
        double pdern = ((a2 + a3 * x) * (*pder) - a4 * pderl + a3 * (*poly)) / a1;
        psave = polyl;
        pdsave = pderl;
        polyl = *poly;
        *poly = polyn;
        pderl = *pder;
        *pder = pdern;
    }
    *polym1 = polyl;
    *pderm1 = pderl;
    *polym2 = psave;
    *pderm2 = pdsave;
}

/* core/speclib.f:371-393 GAMMAF */
static double gammaf(double x)
{
    double pi = 4.0 * atan(1.0);
    double g = 1.0;
    if (x == -0.5) g = -2.0 * sqrt(pi);
    if (x == 0.5) g = sqrt(pi);
    if (x == 1.0) g = 1.0;
    if (x == 2.0) g = 1.0;
    if (x == 1.5) g = sqrt(pi) / 2.0;
    if (x == 2.5) g = 1.5 * sqrt(pi) / 2.0;
    if (x == 3.5) g = 0.5 * (2.5 * (1.5 * sqrt(pi)));
    if (x == 3.0) g = 2.0;
    if (x == 4.0) g = 6.0;
    if (x == 5.0) g = 24.0;
    if (x == 6.0) g = 120.0;
    return g;
}

/* core/speclib.f:395-419 PNORMJ */
static double pnormj(int n, double alpha, double beta)
{
    double dn = (double)n;
    double cst = alpha + beta + 1.0;
    double prod;
    if (n <= 1) {
        prod = gammaf(dn + alpha) * gammaf(dn + beta);
        prod = prod / (gammaf(dn) * gammaf(dn + alpha + beta));
        return prod * pow(2.0, cst) / (2.0 * dn + cst);
    }
    prod = gammaf(alpha + 1.0) * gammaf(beta + 1.0);
    prod = prod / (2.0 * (1.0 + cst) * gammaf(cst + 1.0));
    prod = prod * (1.0 + alpha) * (2.0 + alpha);
    prod = prod * (1.0 + beta) * (2.0 + beta);
    for (int i = 3; i <= n; i++) {
        double dindx = (double)i;
        double frac = (dindx + alpha) * (dindx + beta) / (dindx * (dindx + alpha + beta));
        prod = prod * frac;
    }
    return prod * pow(2.0, cst) / (2.0 * dn + cst);
}

/* core/speclib.f:421-477 JACG */
static void jacg(double *xjac, int np, double alpha, double beta)
{
    const int kstop = 10;
    const double eps = 1.0e-12;
    int n = np - 1;
    double dth = 4.0 * atan(1.0) / (2.0 * (double)n + 2.0);
    double xlast = 0.0;
    for (int j = 1; j <= np; j++) {
        double x;
        if (j == 1) {
            x = cos((2.0 * ((double)j - 1.0) + 1.0) * dth);
        } else {
            double x1 = cos((2.0 * ((double)j - 1.0) + 1.0) * dth);
            double x2 = xlast;
            x = (x1 + x2) / 2.0;
        }
        for (int k = 1; k <= kstop; k++) {
            double p, pd, pm1, pdm1, pm2, pdm2;
            jacobf(&p, &pd, &pm1, &pdm1, &pm2, &pdm2, np, alpha, beta, x);
            double recsum = 0.0;
            int jm = j - 1;
            for (int i = 1; i <= jm; i++) recsum = recsum + 1.0 / (x - xjac[np - i + 1 - 1]);
            double delx = -p / (pd - recsum * p);
            x = x + delx;
            if (fabs(delx) < eps) break;
        }
        xjac[np - j + 1 - 1] = x;
        xlast = x;
    }
    for (int i = 1; i <= np; i++) {
        double xmin = 2.0;
        int jmin = i;
        for (int j = i; j <= np; j++) {
            if (xjac[j - 1] < xmin) {
                xmin = xjac[j - 1];
                jmin = j;
            }
        }
        if (jmin != i) {
            double swap = xjac[i - 1];
            xjac[i - 1] = xjac[jmin - 1];
            xjac[jmin - 1] = swap;
        }
    }
}

/* core/speclib.f:155-205 ZWGJD */
static void zwgjd(double *z, double *w, int np, double alpha, double beta)
{
    int n = np - 1;
    double apb = alpha + beta;
    if (np == 1) {
        z[0] = (beta - alpha) / (apb + 2.0);
        w[0] = gammaf(alpha + 1.0) * gammaf(beta + 1.0) / gammaf(apb + 2.0) * pow(2.0, apb + 1.0);
        return;
    }
    jacg(z, np, alpha, beta);
    int np1 = n + 1, np2 = n + 2;
    double dnp1 = (double)np1, dnp2 = (double)np2;
    double fac1 = dnp1 + alpha + beta + 1.0;
    double fac2 = fac1 + dnp1;
    double fac3 = fac2 + 1.0;
    double fnorm = pnormj(np1, alpha, beta);
    double rcoef = (fnorm * fac2 * fac3) / (2.0 * fac1 * dnp2);
    for (int i = 0; i < np; i++) {
        double p, pd, pm1, pdm1, pm2, pdm2;
        jacobf(&p, &pd, &pm1, &pdm1, &pm2, &pdm2, np2, alpha, beta, z[i]);
        w[i] = -rcoef / (p * pdm1);
    }
}

/* core/speclib.f:283-325 ENDW1 / :327-369 ENDW2 (which = 1 or 2) */
static double endw(int which, int n, double alpha, double beta)
{
    double apb = alpha + beta;
    double f1, f2, f3 = 0.0, fint1, fint2;
    if (n == 0) return 0.0;
    if (which == 1)
        f1 = gammaf(alpha + 2.0) * gammaf(beta + 1.0) / gammaf(apb + 3.0);
    else
        f1 = gammaf(alpha + 1.0) * gammaf(beta + 2.0) / gammaf(apb + 3.0);
    f1 = f1 * (apb + 2.0) * pow(2.0, apb + 2.0) / 2.0;
    if (n == 1) return f1;
    if (which == 1)
        fint1 = gammaf(alpha + 2.0) * gammaf(beta + 1.0) / gammaf(apb + 3.0);
    else
        fint1 = gammaf(alpha + 1.0) * gammaf(beta + 2.0) / gammaf(apb + 3.0);
    fint1 = fint1 * pow(2.0, apb + 2.0);
    fint2 = gammaf(alpha + 2.0) * gammaf(beta + 2.0) / gammaf(apb + 4.0);
    fint2 = fint2 * pow(2.0, apb + 3.0);
    if (which == 1)
        f2 = (-2.0 * (beta + 2.0) * fint1 + (apb + 4.0) * fint2) * (apb + 3.0) / 4.0;
    else
        f2 = (2.0 * (alpha + 2.0) * fint1 - (apb + 4.0) * fint2) * (apb + 3.0) / 4.0;
    if (n == 2) return f2;
    for (int i = 3; i <= n; i++) {
        double di = (double)(i - 1);
        double abn = alpha + beta + di;
        double abnn = abn + di;
        double a1 = -(2.0 * (di + alpha) * (di + beta)) / (abn * abnn * (abnn + 1.0));
        double a2 = (2.0 * (alpha - beta)) / (abnn * (abnn + 2.0));
        double a3 = (2.0 * (abn + 1.0)) / ((abnn + 2.0) * (abnn + 1.0));
        f3 = -(a2 * f2 + a1 * f1) / a3;
        f1 = f2;
        f2 = f3;
    }
    return f3;
}

/* core/speclib.f:107-122 ZWGLL -> :238-281 ZWGLJD with alpha=beta=0 */
void nko_zwgll(double *z, double *w, int np)
{
    double alpha = 0.0, beta = 0.0;
    int n = np - 1, nm1 = n - 1;
    if (nm1 > 0) zwgjd(z + 1, w + 1, nm1, alpha + 1.0, beta + 1.0);
    z[0] = -1.0;
    z[np - 1] = 1.0;
    for (int i = 1; i < np - 1; i++) w[i] = w[i] / (1.0 - z[i] * z[i]);
    double p, pd, pm1, pdm1, pm2, pdm2;
    jacobf(&p, &pd, &pm1, &pdm1, &pm2, &pdm2, n, alpha, beta, z[0]);
    w[0] = endw(1, n, alpha, beta) / (2.0 * pd);
    jacobf(&p, &pd, &pm1, &pdm1, &pm2, &pdm2, n, alpha, beta, z[np - 1]);
    w[np - 1] = endw(2, n, alpha, beta) / (2.0 * pd);
}

/* core/speclib.f:875-905 PNLEG */
static double pnleg(double z, int n)
{
    if (fabs(z) < 1.0e-25) z = 0.0;
    double p1 = 1.0;
    if (n == 0) return p1;
    double p2 = z, p3 = p2;
    for (int k = 1; k <= n - 1; k++) {
        double fk = (double)k;
        p3 = ((2.0 * fk + 1.0) * z * p2 - fk * p1) / (fk + 1.0);
        p1 = p2;
        p2 = p3;
    }
    return p3;
}

/* core/speclib.f:800-833 DGLL.  d, dt are nz x nz column-major: D(I,J)=d[I+nz*J]. */
void nko_dgll(double *d, double *dt, const double *z, int nz)
{
    int n = nz - 1;
    if (nz == 1) {
        d[0] = 0.0;
        return;
    }
    double fn = (double)n;
    double d0 = fn * (fn + 1.0) / 4.0;
    for (int i = 0; i < nz; i++)
        for (int j = 0; j < nz; j++) {
            double v = 0.0;
            if (i != j) v = pnleg(z[i], n) / (pnleg(z[j], n) * (z[i] - z[j]));
            if (i == j && i == 0) v = -d0;
            if (i == j && i == nz - 1) v = d0;
            d[i + nz * j] = v;
            dt[j + nz * i] = v;
        }
}

/* ------------------------------------------------------------------------- */
/* mxm_std.f : small dense matmul with the reference summation order          */
/* ------------------------------------------------------------------------- */

/* core/mxm_wrapper.f:1-87 -> core/mxm_std.f:2-66 mxmf2 -> mxf<n2> (e.g. mxf8
 * :174-191): C(n1,n3) = A(n1,n2) B(n2,n3), column-major, inner sum k=1..n2 left
 * to right, first term is a bare product. */
void nko_mxm(const double *a, int n1, const double *b, int n2, double *c, int n3)
{
    for (int j = 0; j < n3; j++)
        for (int i = 0; i < n1; i++) {
            double s = a[i] * b[n2 * j];
            for (int k = 1; k < n2; k++) s = s + a[i + n1 * k] * b[k + n2 * j];
            c[i + n1 * j] = s;
        }
}

/* core/mxm_std.f:4117-4160 mxma -> mxma2 -> mxa<n2>: C = C + A B, sum order
 * ((c + a1 b1) + a2 b2) + ... */
void nko_mxma(const double *a, int n1, const double *b, int n2, double *c, int n3)
{
    for (int j = 0; j < n3; j++)
        for (int i = 0; i < n1; i++) {
            double s = c[i + n1 * j];
            for (int k = 0; k < n2; k++) s = s + a[i + n1 * k] * b[k + n2 * j];
            c[i + n1 * j] = s;
        }
}

/* ------------------------------------------------------------------------- */
/* Box mesh in genbox element order + lexicographic vertex ids                */
/* ------------------------------------------------------------------------- */

/* tools/genbox (element order x fastest, examples/bp5/genbox.in:16-20) and the
 * corner conventions of core/genxyz.f:1279-1291: xc,yc,zc(8,E) are written in
 * PREPROCESSOR corner order (as stored in .re2), vertex(8,E) in SYMMETRIC
 * (hypercube) order as delivered by .ma2 (core/map2.f:139).  genmap's actual
 * vertex ids are an arbitrary labelling; here vertex id = 1 + lexicographic
 * grid-vertex index, with wrap-around in periodic directions.  per[d]!=0 makes
 * direction d periodic. */
void nko_box_mesh(int nelx, int nely, int nelz, const double *lo, const double *hi,
                  const int *per, double *xc, double *yc, double *zc, int64_t *vertex)
{
    static const int indx[8] = {1, 2, 4, 3, 5, 6, 8, 7}; /* genxyz.f:1291 sym->prex */
    int nvx = per[0] ? nelx : nelx + 1;
    int nvy = per[1] ? nely : nely + 1;
    int nvz = per[2] ? nelz : nelz + 1;
    for (int iz = 0; iz < nelz; iz++)
        for (int iy = 0; iy < nely; iy++)
            for (int ix = 0; ix < nelx; ix++) {
                int64_t e = ix + (int64_t)nelx * (iy + (int64_t)nely * iz);
                for (int k = 0; k < 2; k++)
                    for (int j = 0; j < 2; j++)
                        for (int i = 0; i < 2; i++) {
                            int is = i + 2 * j + 4 * k; /* symmetric corner, 0-based */
                            int ip = indx[is] - 1;      /* preprocessor corner */
                            double x = lo[0] + (hi[0] - lo[0]) * (double)(ix + i) / (double)nelx;
                            double y = lo[1] + (hi[1] - lo[1]) * (double)(iy + j) / (double)nely;
                            double z = lo[2] + (hi[2] - lo[2]) * (double)(iz + k) / (double)nelz;
                            xc[ip + 8 * e] = x;
                            yc[ip + 8 * e] = y;
                            zc[ip + 8 * e] = z;
                            int vx = ix + i, vy = iy + j, vz = iz + k;
                            if (per[0]) vx %= nvx;
                            if (per[1]) vy %= nvy;
                            if (per[2]) vz %= nvz;
                            vertex[is + 8 * e] = 1 + vx + (int64_t)nvx * (vy + (int64_t)nvy * vz);
                        }
            }
}

/* core/genxyz.f:1269-1332 xyzlin: trilinear map of the 8 corners onto the GLL
 * nodes via tensr3 (core/fasts.f:125-162) with the 2-point Lagrange weights that
 * fd_weights_full (core/fast3d.f:1294-1349) returns for nodes (-1,1):
 * J(i,1)=(1-z_i)/2, J(i,2)=(1+z_i)/2.  zg = GLL points (nx). */
void nko_xyzlin(int nx, int64_t nel, const double *zg, const double *xc, const double *yc,
                const double *zc, double *xm1, double *ym1, double *zm1)
{
    int nxyz = nx * nx * nx;
    double *jx = malloc(sizeof(double) * nx * 2);  /* jx(nx,2)  */
    double *jxt = malloc(sizeof(double) * 2 * nx); /* jxt(2,nx) */
    double *v = malloc(sizeof(double) * nxyz);
    double *w = malloc(sizeof(double) * nxyz);
    static const int indx[8] = {1, 2, 4, 3, 5, 6, 8, 7};
    for (int i = 0; i < nx; i++) {
        double c0 = (1.0 - zg[i]) / 2.0, c1 = (1.0 + zg[i]) / 2.0;
        jxt[0 + 2 * i] = c0;
        jxt[1 + 2 * i] = c1;
        jx[i] = c0;
        jx[i + nx] = c1;
    }
    for (int64_t e = 0; e < nel; e++) {
        const double *src[3] = {xc + 8 * e, yc + 8 * e, zc + 8 * e};
        double *dst[3] = {xm1 + nxyz * e, ym1 + nxyz * e, zm1 + nxyz * e};
        for (int c = 0; c < 3; c++) {
            double cb[8];
            for (int ix = 0; ix < 8; ix++) cb[ix] = src[c][indx[ix] - 1];
            /* tensr3(v,nv=nx,u=cb,nu=2,A=jx,Bt=jxt,Ct=jxt,w) */
            nko_mxm(jx, nx, cb, 2, v, 4);
            for (int iz = 0; iz < 2; iz++) nko_mxm(v + iz * 2 * nx, nx, jxt, 2, w + iz * nx * nx, nx);
            nko_mxm(w, nx * nx, jxt, 2, dst[c], nx);
        }
    }
    free(jx);
    free(jxt);
    free(v);
    free(w);
}

/* core/navier5.f:2702-2718 rescale_x (examples/bp5/bp5.usr:40-44) */
void nko_rescale_x(double *x, int64_t n, double x0, double x1)
{
    double xmin = x[0], xmax = x[0];
    for (int64_t i = 0; i < n; i++) {
        if (x[i] < xmin) xmin = x[i];
        if (x[i] > xmax) xmax = x[i];
    }
    if (xmax <= xmin) return;
    double scale = (x1 - x0) / (xmax - xmin);
    for (int64_t i = 0; i < n; i++) x[i] = x0 + scale * (x[i] - xmin);
}

/* ------------------------------------------------------------------------- */
/* Geometric factors                                                          */
/* ------------------------------------------------------------------------- */

/* local gradient: examples/bp5/bp5.usr:56-75 loc_grad3 == core/coef.f:879-927
 * xyzrst (same three mxm calls).  d = D(nx,nx), dt = D^T. */
static void loc_grad3(double *ur, double *us, double *ut, const double *u, int nx,
                      const double *d, const double *dt)
{
    int m1 = nx, m2 = nx * nx;
    nko_mxm(d, m1, u, m1, ur, m2);
    for (int k = 0; k < nx; k++) nko_mxm(u + k * m2, m1, dt, m1, us + k * m2, m1);
    nko_mxm(u, m2, dt, m1, ut, m1);
}

/* examples/bp5/bp5.usr:77-95 loc_grad3t */
static void loc_grad3t(double *u, const double *ur, const double *us, const double *ut, int nx,
                       const double *d, const double *dt)
{
    int m1 = nx, m2 = nx * nx;
    nko_mxm(dt, m1, ur, m1, u, m2);
    for (int k = 0; k < nx; k++) nko_mxma(us + k * m2, m1, d, m1, u + k * m2, m1);
    nko_mxma(ut, m2, d, m1, u, m1);
}

/* core/coef.f:555-631 glmapm1 (3-D branch) + :633-784 geodat1 (3-D, non-axisymmetric).
 * Outputs g1..g6 in the core order rr,ss,tt,rs,rt,st (core/GEOM:42-48), bm1, jacm1.
 * w3 = w3m1(nx,nx,nx) (core/coef.f:263-267). */
void nko_geom_core(int nx, int64_t nel, const double *d, const double *dt, const double *w3,
                   const double *xm1, const double *ym1, const double *zm1, double *g1, double *g2,
                   double *g3, double *g4, double *g5, double *g6, double *bm1, double *jacm1)
{
    int n = nx * nx * nx;
    double *buf = malloc(sizeof(double) * 9 * n);
    double *xr = buf, *xs = buf + n, *xt = buf + 2 * n, *yr = buf + 3 * n, *ys = buf + 4 * n,
           *yt = buf + 5 * n, *zr = buf + 6 * n, *zs = buf + 7 * n, *zt = buf + 8 * n;
    for (int64_t e = 0; e < nel; e++) {
        loc_grad3(xr, xs, xt, xm1 + n * e, nx, d, dt);
        loc_grad3(yr, ys, yt, ym1 + n * e, nx, d, dt);
        loc_grad3(zr, zs, zt, zm1 + n * e, nx, d, dt);
        for (int i = 0; i < n; i++) {
            /* coef.f:610-615: rzero + addcol4 x3 + subcol4 x3 */
            double jac = 0.0;
            jac = jac + xr[i] * ys[i] * zt[i];
            jac = jac + xt[i] * yr[i] * zs[i];
            jac = jac + xs[i] * yt[i] * zr[i];
            jac = jac - xr[i] * yt[i] * zs[i];
            jac = jac - xs[i] * yr[i] * zt[i];
            jac = jac - xt[i] * ys[i] * zr[i];
            /* coef.f:616-624 ascol5: a = b*c - d*e */
            double rx = ys[i] * zt[i] - yt[i] * zs[i];
            double ry = xt[i] * zs[i] - xs[i] * zt[i];
            double rz = xs[i] * yt[i] - xt[i] * ys[i];
            double sx = yt[i] * zr[i] - yr[i] * zt[i];
            double sy = xr[i] * zt[i] - xt[i] * zr[i];
            double sz = xt[i] * yr[i] - xr[i] * yt[i];
            double tx = yr[i] * zs[i] - ys[i] * zr[i];
            double ty = xs[i] * zr[i] - xr[i] * zs[i];
            double tz = xr[i] * ys[i] - xs[i] * yr[i];
            double wj = 1.0 / jac; /* geodat1: invers2(wj,jacm1) */
            int64_t q = i + (int64_t)n * e;
            /* vdot3 (math.f): a = b1*c1 + b2*c2 + b3*c3 ; col2(wj) ; col2(w3m1) */
            g1[q] = (rx * rx + ry * ry + rz * rz) * wj * w3[i];
            g2[q] = (sx * sx + sy * sy + sz * sz) * wj * w3[i];
            g3[q] = (tx * tx + ty * ty + tz * tz) * wj * w3[i];
            g4[q] = (rx * sx + ry * sy + rz * sz) * wj * w3[i];
            g5[q] = (rx * tx + ry * ty + rz * tz) * wj * w3[i];
            g6[q] = (sx * tx + sy * ty + sz * tz) * wj * w3[i];
            bm1[q] = jac * w3[i]; /* col3(bm1,jacm1,w3m1) */
            jacm1[q] = jac;
        }
    }
    free(buf);
}

/* examples/bp5/bp5.usr:623-699 geodatstd: gf(6,nxyz,E) interleaved, order
 * rr,rs,rt,ss,st,tt, scaled by w*J. */
void nko_geodatstd(int nx, int64_t nel, const double *d, const double *dt, const double *w3,
                   const double *xm1, const double *ym1, const double *zm1, double *gf)
{
    int n = nx * nx * nx;
    double *buf = malloc(sizeof(double) * 9 * n);
    double *xr = buf, *xs = buf + n, *xt = buf + 2 * n, *yr = buf + 3 * n, *ys = buf + 4 * n,
           *yt = buf + 5 * n, *zr = buf + 6 * n, *zs = buf + 7 * n, *zt = buf + 8 * n;
    for (int64_t e = 0; e < nel; e++) {
        loc_grad3(xr, xs, xt, xm1 + n * e, nx, d, dt);
        loc_grad3(yr, ys, yt, ym1 + n * e, nx, d, dt);
        loc_grad3(zr, zs, zt, zm1 + n * e, nx, d, dt);
        for (int i = 0; i < n; i++) {
            double jacmq = xr[i] * (ys[i] * zt[i] - yt[i] * zs[i]) -
                           xs[i] * (yr[i] * zt[i] - yt[i] * zr[i]) +
                           xt[i] * (yr[i] * zs[i] - ys[i] * zr[i]);
            double a11 = xr[i], a12 = xs[i], a13 = xt[i];
            double a21 = yr[i], a22 = ys[i], a23 = yt[i];
            double a31 = zr[i], a32 = zs[i], a33 = zt[i];
            double g11 = (a22 * a33 - a23 * a32) / jacmq;
            double g12 = (a13 * a32 - a33 * a12) / jacmq;
            double g13 = (a12 * a23 - a22 * a13) / jacmq;
            double g21 = (a23 * a31 - a21 * a33) / jacmq;
            double g22 = (a11 * a33 - a31 * a13) / jacmq;
            double g23 = (a13 * a21 - a23 * a11) / jacmq;
            double g31 = (a21 * a32 - a22 * a31) / jacmq;
            double g32 = (a12 * a31 - a32 * a11) / jacmq;
            double g33 = (a11 * a22 - a21 * a12) / jacmq;
            double scale = w3[i] * jacmq;
            double *g = gf + 6 * ((int64_t)i + (int64_t)n * e);
            g[0] = scale * (g11 * g11 + g12 * g12 + g13 * g13);
            g[1] = scale * (g11 * g21 + g12 * g22 + g13 * g23);
            g[2] = scale * (g11 * g31 + g12 * g32 + g13 * g33);
            g[3] = scale * (g21 * g21 + g22 * g22 + g23 * g23);
            g[4] = scale * (g21 * g31 + g22 * g32 + g23 * g33);
            g[5] = scale * (g31 * g31 + g32 * g32 + g33 * g33);
        }
    }
    free(buf);
}

/* ------------------------------------------------------------------------- */
/* Global node numbering (integer, bit-exact target)                          */
/* ------------------------------------------------------------------------- */

typedef struct {
    int64_t key[3];
    int64_t src;
    int64_t rank;
} tup_t;

static int g_nkey;
static int tup_cmp(const void *pa, const void *pb)
{
    const tup_t *a = pa, *b = pb;
    for (int k = 0; k < g_nkey; k++) {
        if (a->key[k] < b->key[k]) return -1;
        if (a->key[k] > b->key[k]) return 1;
    }
    return 0;
}

/* core/navier8.f:1934-2002 gbtuple_rank8, all ranks emulated at once.  Tuples go
 * to processor mod(key1,np) (:1964); there they are sorted lexicographically and
 * unique tuples ranked 1..nu (i8rank_vecn :1790-1830; the heap sort's order among
 * equal tuples does not affect ranks, so qsort is used); ranks are offset by the
 * running sum of nu over lower processors (:1986-1991).  Result in t[i].rank. */
static void gbtuple_rank(tup_t *t, int64_t n, int nkey, int np)
{
    int64_t *start = calloc((size_t)np + 1, sizeof(int64_t));
    for (int64_t i = 0; i < n; i++) start[1 + (int)(t[i].key[0] % np)]++;
    for (int p = 0; p < np; p++) start[p + 1] += start[p];
    tup_t *b = malloc(sizeof(tup_t) * (size_t)(n > 0 ? n : 1));
    int64_t *fill = malloc(sizeof(int64_t) * (size_t)np);
    for (int p = 0; p < np; p++) fill[p] = start[p];
    for (int64_t i = 0; i < n; i++) b[fill[t[i].key[0] % np]++] = t[i];
    int64_t nu_prior = 0;
    g_nkey = nkey;
    for (int p = 0; p < np; p++) {
        int64_t ni = start[p + 1] - start[p];
        tup_t *bp = b + start[p];
        if (ni == 0) continue;
        qsort(bp, (size_t)ni, sizeof(tup_t), tup_cmp);
        int64_t nn = 1;
        bp[0].rank = nn + nu_prior;
        for (int64_t i = 1; i < ni; i++) {
            if (tup_cmp(&bp[i - 1], &bp[i]) != 0) nn++;
            bp[i].rank = nn + nu_prior;
        }
        nu_prior += nn;
    }
    for (int64_t i = 0; i < n; i++) t[b[i].src].rank = b[i].rank;
    free(b);
    free(fill);
    free(start);
}

/* core/navier8.f:1131-1180 i8rank (heap sort, index form), literal for n<=8. */
static void i8rank(const int64_t *a, int *ind, int n)
{
    /* 1-based emulation */
    if (n <= 1) return;
    for (int j = 1; j <= n; j++) ind[j - 1] = j;
    int l = n / 2 + 1, ir = n, indx, i, j;
    int64_t q;
    for (;;) {
        if (l > 1) {
            l--;
            indx = ind[l - 1];
            q = a[indx - 1];
        } else {
            indx = ind[ir - 1];
            q = a[indx - 1];
            ind[ir - 1] = ind[0];
            ir--;
            if (ir == 1) {
                ind[0] = indx;
                return;
            }
        }
        i = l;
        j = l + l;
        while (j <= ir) {
            if (j < ir) {
                if (a[ind[j - 1] - 1] < a[ind[j] - 1]) j++;
            }
            if (q < a[ind[j - 1] - 1]) {
                ind[i - 1] = ind[j - 1];
                i = j;
                j = j + j;
            } else {
                j = ir + 1;
            }
        }
        ind[i - 1] = indx;
    }
}

static int i64_cmp(const void *a, const void *b)
{
    int64_t x = *(const int64_t *)a, y = *(const int64_t *)b;
    return (x > y) - (x < y);
}

/* core/navier8.f:2004-2360 setvert3d (ifcenter=.false.) for ALL elements of the
 * mesh at once, emulating an np-rank run (the element->rank map does not
 * influence the ids, only np does, through gbtuple_rank8's mod-np bucketing).
 * vertex(8,nel) symmetric order; glo_num(nx^3,nel); returns ngv. */
int64_t nko_setvert3d(int64_t *glo_num, int nx, int64_t nel, const int64_t *vertex, int np)
{
    static const int icface[6][4] = {{1, 3, 5, 7}, {2, 4, 6, 8}, {1, 2, 5, 6},
                                     {3, 4, 7, 8}, {1, 2, 3, 4}, {5, 6, 7, 8}}; /* core/TOPOL:37-41 */
    int ny = nx, nz = nx;
    int64_t nxyz = (int64_t)nx * ny * nz;
    int64_t ngvv = 0;
    for (int64_t i = 0; i < 8 * nel; i++)
        if (vertex[i] > ngvv) ngvv = vertex[i]; /* :2050 i8glmax */
    memset(glo_num, 0, sizeof(int64_t) * (size_t)(nxyz * nel));
    /* :2052-2062 vertices */
    for (int64_t e = 0; e < nel; e++)
        for (int k = 0; k < 2; k++)
            for (int j = 0; j < 2; j++)
                for (int i = 0; i < 2; i++) {
                    int64_t il = (nx - 1) * i + (int64_t)nx * (nx - 1) * j + (int64_t)nx * nx * (nx - 1) * k;
                    glo_num[il + nxyz * e] = vertex[i + 2 * j + 4 * k + 8 * e];
                }
    int64_t ngv = ngvv;
    if (nx == 2) return ngv;

    /* :2071-2092 edge labels by sorted bounding vertices */
    tup_t *et = malloc(sizeof(tup_t) * (size_t)(12 * nel));
    for (int64_t e = 0; e < nel; e++) {
        const int64_t *v = vertex + 8 * e;
        for (int c = 0; c < 2; c++)
            for (int b = 0; b < 2; b++) {
                int64_t p0[3], p1[3];
                /* r-edge (d=1): j=b,k=c ; s-edge (d=2): i=b,k=c ; t-edge (d=3): i=b,j=c */
                p0[0] = v[0 + 2 * b + 4 * c];
                p1[0] = v[1 + 2 * b + 4 * c];
                p0[1] = v[b + 2 * 0 + 4 * c];
                p1[1] = v[b + 2 * 1 + 4 * c];
                p0[2] = v[b + 2 * c + 4 * 0];
                p1[2] = v[b + 2 * c + 4 * 1];
                for (int d = 0; d < 3; d++) {
                    int64_t idx = 12 * e + (b + 2 * c + 4 * d);
                    int64_t lo = p0[d], hi = p1[d];
                    if (lo > hi) {
                        int64_t s = lo;
                        lo = hi;
                        hi = s;
                    }
                    et[idx].key[0] = lo;
                    et[idx].key[1] = hi;
                    et[idx].key[2] = 0;
                    et[idx].src = idx;
                }
            }
    }
    gbtuple_rank(et, 12 * nel, 2, np); /* :2098 */
    int64_t n_unique_edges = 0;
    for (int64_t i = 0; i < 12 * nel; i++)
        if (et[i].rank > n_unique_edges) n_unique_edges = et[i].rank; /* :2102 */
    int64_t n_on_edge = nx - 2;
    int64_t ngve = n_unique_edges * n_on_edge;
    for (int64_t e = 0; e < nel; e++) {
        int iedg = 0;
        int64_t *g = glo_num + nxyz * e; /* g[idx-1] <-> glo_num(idx + nxyz*(e-1)) */
        /* :2110-2127 edges 1-4 */
        for (int k = 0; k < 2; k++)
            for (int j = 0; j < 2; j++) {
                int64_t igv = ngv + n_on_edge * (et[12 * e + iedg].rank - 1);
                int64_t i0 = (int64_t)nx * (nx - 1) * j + (int64_t)nx * nx * (nx - 1) * k;
                if (g[i0 + 1 - 1] < g[i0 + nx - 1]) {
                    for (int i = 2; i <= nx - 1; i++) g[i0 + i - 1] = igv + i - 1;
                } else {
                    for (int i = 2; i <= nx - 1; i++) g[i0 + i - 1] = igv + 1 + n_on_edge - (i - 1);
                }
                iedg++;
            }
        /* :2129-2146 edges 5-8 */
        for (int k = 0; k < 2; k++)
            for (int i = 0; i < 2; i++) {
                int64_t igv = ngv + n_on_edge * (et[12 * e + iedg].rank - 1);
                int64_t i0 = 1 + (nx - 1) * i + (int64_t)nx * nx * (nx - 1) * k;
                if (g[i0 - 1] < g[i0 + (int64_t)nx * (nx - 1) - 1]) {
                    for (int j = 2; j <= nx - 1; j++) g[i0 + (j - 1) * nx - 1] = igv + j - 1;
                } else {
                    for (int j = 2; j <= nx - 1; j++)
                        g[i0 + (j - 1) * nx - 1] = igv + 1 + n_on_edge - (j - 1);
                }
                iedg++;
            }
        /* :2148-2165 edges 9-12 */
        for (int j = 0; j < 2; j++)
            for (int i = 0; i < 2; i++) {
                int64_t igv = ngv + n_on_edge * (et[12 * e + iedg].rank - 1);
                int64_t i0 = 1 + (nx - 1) * i + (int64_t)nx * (nx - 1) * j;
                if (g[i0 - 1] < g[i0 + (int64_t)nx * nx * (nx - 1) - 1]) {
                    for (int k = 2; k <= nx - 1; k++) g[i0 + (int64_t)(k - 1) * nx * nx - 1] = igv + k - 1;
                } else {
                    for (int k = 2; k <= nx - 1; k++)
                        g[i0 + (int64_t)(k - 1) * nx * nx - 1] = igv + 1 + n_on_edge - (k - 1);
                }
                iedg++;
            }
    }
    ngv = ngv + ngve;
    free(et);

    /* :2189-2210 faces by their 3 smallest vertices */
    tup_t *ft = malloc(sizeof(tup_t) * (size_t)(6 * nel));
    for (int64_t e = 0; e < nel; e++)
        for (int ifac = 0; ifac < 6; ifac++) {
            int64_t facet[4];
            for (int icrn = 0; icrn < 4; icrn++) facet[icrn] = vertex[icface[ifac][icrn] - 1 + 8 * e];
            qsort(facet, 4, sizeof(int64_t), i64_cmp); /* :2197 i8sort */
            int64_t idx = ifac + 6 * e;
            ft[idx].key[0] = facet[0];
            ft[idx].key[1] = facet[1];
            ft[idx].key[2] = facet[2];
            ft[idx].src = idx;
        }
    gbtuple_rank(ft, 6 * nel, 3, np); /* :2206 */
    int64_t n_unique_faces = 0;
    for (int64_t i = 0; i < 6 * nel; i++)
        if (ft[i].rank > n_unique_faces) n_unique_faces = ft[i].rank;

    /* core/connect1.f:546-620 dsset: skpdat(1..6, face) */
    int skp[6][6] = {
        {1, nx * (ny - 1) + 1, nx, 1, ny * (nz - 1) + 1, ny},
        {1 + (nx - 1), nx * (ny - 1) + 1 + (nx - 1), nx, 1, ny * (nz - 1) + 1, ny},
        {1, nx, 1, 1, ny * (nz - 1) + 1, ny},
        {1 + nx * (ny - 1), nx + nx * (ny - 1), 1, 1, ny * (nz - 1) + 1, ny},
        {1, nx, 1, 1, ny, 1},
        {1 + nx * ny * (nz - 1), nx + nx * ny * (nz - 1), 1, 1, ny, 1}};
    int64_t n_on_face = (int64_t)(nx - 2) * (ny - 2);
    int64_t ngvs = n_unique_faces * n_on_face;
    int nxx = nx * nx;
    for (int64_t e = 0; e < nel; e++) {
        int64_t *g = glo_num + nxyz * e;
        for (int iface = 0; iface < 6; iface++) {
            int i0 = skp[iface][0], i1 = skp[iface][1], is = skp[iface][2];
            int j0 = skp[iface][3], j1 = skp[iface][4], js = skp[iface][5];
            int64_t gvf[4];
            int ind[4];
            gvf[0] = g[i0 + nx * (j0 - 1) - 1];
            gvf[1] = g[i1 + nx * (j0 - 1) - 1];
            gvf[2] = g[i0 + nx * (j1 - 1) - 1];
            gvf[3] = g[i1 + nx * (j1 - 1) - 1];
            i8rank(gvf, ind, 4); /* :2241 */
            int ifij = 0, idir = 1, jdir = 1;
            if (ind[0] == 1) {
                idir = 1, jdir = 1;
                if (gvf[1] < gvf[2]) ifij = 1;
            } else if (ind[0] == 2) {
                idir = -1, jdir = 1;
                if (gvf[0] < gvf[3]) ifij = 1;
            } else if (ind[0] == 3) {
                idir = 1, jdir = -1;
                if (gvf[3] < gvf[0]) ifij = 1;
            } else if (ind[0] == 4) {
                idir = -1, jdir = -1;
                if (gvf[2] < gvf[1]) ifij = 1;
            }
            if (idir < 0) {
                int it = i0;
                i0 = i1;
                i1 = it;
                is = -is;
            }
            if (jdir < 0) {
                int jt = j0;
                j0 = j1;
                j1 = jt;
                js = -js;
            }
            int64_t ig0 = ngv + n_on_face * (ft[iface + 6 * e].rank - 1);
            int k = 0;
            int64_t l = 0;
            if (ifij) {
                for (int j = j0; (js > 0) ? (j <= j1) : (j >= j1); j += js)
                    for (int i = i0; (is > 0) ? (i <= i1) : (i >= i1); i += is) {
                        k++;
                        if (k > nx && k < nxx - nx && (k % nx) != 1 && (k % nx) != 0) {
                            l++;
                            g[i + nx * (j - 1) - 1] = l + ig0;
                        }
                    }
            } else {
                for (int i = i0; (is > 0) ? (i <= i1) : (i >= i1); i += is)
                    for (int j = j0; (js > 0) ? (j <= j1) : (j >= j1); j += js) {
                        k++;
                        if (k > nx && k < nxx - nx && (k % nx) != 1 && (k % nx) != 0) {
                            l++;
                            g[i + nx * (j - 1) - 1] = l + ig0;
                        }
                    }
            }
        }
    }
    ngv = ngv + ngvs;
    free(ft);
    /* :2336-2347 interiors are 0 (already zeroed; faces/edges/vertices overwrote) */
    return ngv;
}

/* core/navier8.f:2571-2624 check_p_bc: pflag[3*e+d] != 0 when element e carries
 * 'p  ' on both faces of direction d (single-element periodicity). */
void nko_check_p_bc(int64_t *glo_num, int nx, int64_t nel, const int *pflag)
{
    int64_t nxyz = (int64_t)nx * nx * nx;
    for (int64_t e = 0; e < nel; e++) {
        int64_t *g = glo_num + nxyz * e;
        for (int d = 0; d < 3; d++) {
            if (!pflag[3 * e + d]) continue;
            int s = (d == 0) ? 1 : (d == 1 ? nx : nx * nx);
            for (int b = 0; b < nx; b++)
                for (int a = 0; a < nx; a++) {
                    int64_t base;
                    if (d == 0)
                        base = (int64_t)nx * (a + (int64_t)nx * b);
                    else if (d == 1)
                        base = a + (int64_t)nx * nx * b;
                    else
                        base = a + (int64_t)nx * b;
                    int64_t lo = base, hi = base + (int64_t)(nx - 1) * s;
                    int64_t gmn = g[lo] < g[hi] ? g[lo] : g[hi];
                    g[lo] = gmn;
                    g[hi] = gmn;
                }
        }
    }
}

/* ------------------------------------------------------------------------- */
/* gather-scatter (gslib v1.0.9 gs_op semantics; library not in the tree)     */
/* ------------------------------------------------------------------------- */

typedef struct {
    int64_t id;
    int64_t idx;
} idp_t;
static int idp_cmp(const void *pa, const void *pb)
{
    const idp_t *a = pa, *b = pb;
    if (a->id != b->id) return (a->id > b->id) - (a->id < b->id);
    return (a->idx > b->idx) - (a->idx < b->idx);
}

/* gslib v1.0.9 (3rd_party/gslib/install:5) gs_op as used at core/dssum.f:79 and
 * :140-158: every set of entries sharing one non-zero id is replaced by its
 * sum / product / min / max (op = 1,2,3,4 as in dsop core/dssum.f:100-161);
 * id 0 entries are untouched.  Combination order: ascending local index. */
void nko_gs_op(double *u, const int64_t *id, int64_t n, int op)
{
    int64_t m = 0;
    for (int64_t i = 0; i < n; i++)
        if (id[i] != 0) m++;
    idp_t *p = malloc(sizeof(idp_t) * (size_t)(m > 0 ? m : 1));
    m = 0;
    for (int64_t i = 0; i < n; i++)
        if (id[i] != 0) {
            p[m].id = id[i];
            p[m].idx = i;
            m++;
        }
    qsort(p, (size_t)m, sizeof(idp_t), idp_cmp);
    int64_t s = 0;
    while (s < m) {
        int64_t t = s + 1;
        while (t < m && p[t].id == p[s].id) t++;
        double v = u[p[s].idx];
        for (int64_t q = s + 1; q < t; q++) {
            double w = u[p[q].idx];
            if (op == 1)
                v = v + w;
            else if (op == 2)
                v = v * w;
            else if (op == 3)
                v = (w < v) ? w : v;
            else
                v = (w > v) ? w : v;
        }
        for (int64_t q = s; q < t; q++) u[p[q].idx] = v;
        s = t;
    }
    free(p);
}

/* ------------------------------------------------------------------------- */
/* Inputs: random field                                                       */
/* ------------------------------------------------------------------------- */

/* core/navier5.f:2650-2684 ran1 (NR 2nd ed. p.271) with its SAVEd state. */
typedef struct {
    int iv[32];
    int iy;
} ran1_state;

static double ran1(int *idum, ran1_state *st)
{
    const int ia = 16807, im = 2147483647, iq = 127773, ir = 2836, ntab = 32;
    const int ndiv = 1 + (im - 1) / ntab;
    const double am = 1.0 / im, eps = 1.2e-7, rnmx = 1.0 - eps;
    int j, k;
    if (*idum <= 0 || st->iy == 0) {
        *idum = (-*idum > 1) ? -*idum : 1;
        for (j = ntab + 8; j >= 1; j--) {
            k = *idum / iq;
            *idum = ia * (*idum - k * iq) - ir * k;
            if (*idum < 0) *idum = *idum + im;
            if (j <= ntab) st->iv[j - 1] = *idum;
        }
        st->iy = st->iv[0];
    }
    k = *idum / iq;
    *idum = ia * (*idum - k * iq) - ir * k;
    if (*idum < 0) *idum = *idum + im;
    j = 1 + st->iy / ndiv;
    st->iy = st->iv[j - 1];
    st->iv[j - 1] = *idum;
    double r = am * st->iy;
    return r < rnmx ? r : rnmx;
}

/* core/navier5.f:2687-2698 rand_fld_h1 WITHOUT the trailing dsavg (the caller
 * applies dssum*vmult, core/ic.f:1871-1895).  First call of a fresh process:
 * iy=0 so ran1 re-seeds with idum=max(-n,1)=1 (navier5.f:2665-2666). */
void nko_rand_fld(double *x, int64_t n)
{
    ran1_state st;
    memset(&st, 0, sizeof(st));
    int id = (int)n;
    for (int64_t i = 0; i < n; i++) x[i] = ran1(&id, &st);
}

/* ------------------------------------------------------------------------- */
/* Operators                                                                  */
/* ------------------------------------------------------------------------- */

/* examples/bp5/bp5.usr:1278-1307 ax_e_bp5 + :1309-1341 axhm1_bp5.  gf(6,nxyz,E)
 * order rr,rs,rt,ss,st,tt.  Returns pap = sum p*Ap accumulated element by
 * element, point by point (:1334-1336). */
double nko_axhm1_bp5(double *ap, const double *p, const double *gf, int nx, int64_t nel,
                     const double *d, const double *dt)
{
    int n = nx * nx * nx;
    double *ur = malloc(sizeof(double) * 3 * n), *us = ur + n, *ut = ur + 2 * n;
    double pap = 0.0;
    for (int64_t e = 0; e < nel; e++) {
        const double *u = p + (int64_t)n * e;
        const double *g = gf + 6 * (int64_t)n * e;
        double *w = ap + (int64_t)n * e;
        loc_grad3(ur, us, ut, u, nx, d, dt);
        for (int i = 0; i < n; i++) {
            double wr = g[6 * i + 0] * ur[i] + g[6 * i + 1] * us[i] + g[6 * i + 2] * ut[i];
            double ws = g[6 * i + 1] * ur[i] + g[6 * i + 3] * us[i] + g[6 * i + 4] * ut[i];
            double wt = g[6 * i + 2] * ur[i] + g[6 * i + 4] * us[i] + g[6 * i + 5] * ut[i];
            ur[i] = wr;
            us[i] = ws;
            ut[i] = wt;
        }
        loc_grad3t(w, ur, us, ut, nx, d, dt);
        for (int i = 0; i < n; i++) pap = pap + u[i] * w[i];
    }
    free(ur);
    return pap;
}

/* core/hmholtz.f:72-259 axhelm, 3-D, general (non-fast) branch :191-217 with
 * ifdfrm(e) per element (NULL = all deformed, the param(59)=1 default), then
 * :225 addcol4(au,helm2,bm1,u) when ifh2.  g1..g6 core order rr,ss,tt,rs,rt,st. */
void nko_axhelm(double *au, const double *u, const double *h1, const double *h2, int ifh2, int nx,
                int64_t nel, const double *d, const double *dt, const double *g1, const double *g2,
                const double *g3, const double *g4, const double *g5, const double *g6,
                const double *bm1, const int *ifdfrm)
{
    int n = nx * nx * nx, nxy = nx * nx;
    double *buf = malloc(sizeof(double) * 9 * n);
    double *dudr = buf, *duds = buf + n, *dudt = buf + 2 * n, *tmp1 = buf + 3 * n,
           *tmp2 = buf + 4 * n, *tmp3 = buf + 5 * n, *tm1 = buf + 6 * n, *tm2 = buf + 7 * n,
           *tm3 = buf + 8 * n;
    for (int64_t q = 0; q < (int64_t)n * nel; q++) au[q] = 0.0; /* :125 rzero */
    for (int64_t e = 0; e < nel; e++) {
        int64_t o = (int64_t)n * e;
        const double *ue = u + o;
        nko_mxm(d, nx, ue, nx, dudr, nxy);                                              /* :191 */
        for (int iz = 0; iz < nx; iz++) nko_mxm(ue + iz * nxy, nx, dt, nx, duds + iz * nxy, nx); /* :193 */
        nko_mxm(ue, nxy, dt, nx, dudt, nx);                                             /* :195 */
        int dfrm = ifdfrm ? ifdfrm[e] : 1;
        for (int i = 0; i < n; i++) {
            double t1 = dudr[i] * g1[o + i]; /* col3 :196-198 */
            double t2 = duds[i] * g2[o + i];
            double t3 = dudt[i] * g3[o + i];
            if (dfrm) { /* addcol3 :200-205 */
                t1 = t1 + duds[i] * g4[o + i];
                t1 = t1 + dudt[i] * g5[o + i];
                t2 = t2 + dudr[i] * g4[o + i];
                t2 = t2 + dudt[i] * g6[o + i];
                t3 = t3 + dudr[i] * g5[o + i];
                t3 = t3 + duds[i] * g6[o + i];
            }
            tmp1[i] = t1 * h1[o + i]; /* col2 :207-209 */
            tmp2[i] = t2 * h1[o + i];
            tmp3[i] = t3 * h1[o + i];
        }
        nko_mxm(dt, nx, tmp1, nx, tm1, nxy);                                            /* :210 */
        for (int iz = 0; iz < nx; iz++) nko_mxm(tmp2 + iz * nxy, nx, d, nx, tm2 + iz * nxy, nx); /* :212 */
        nko_mxm(tmp3, nxy, d, nx, tm3, nx);                                             /* :214 */
        for (int i = 0; i < n; i++) { /* add2 x3 :215-217 */
            double a = au[o + i];
            a = a + tm1[i];
            a = a + tm2[i];
            a = a + tm3[i];
            au[o + i] = a;
        }
    }
    if (ifh2) /* :225 addcol4: a = a + b*c*d */
        for (int64_t q = 0; q < (int64_t)n * nel; q++) au[q] = au[q] + h2[q] * bm1[q] * u[q];
    free(buf);
}

/* core/hmholtz.f:380-524 setprec (3-D, non-axisymmetric) WITHOUT the trailing
 * dssum + invcol1 (:520-521), which the caller applies.  dxt = DXTM1 = D^T:
 * DXTM1(ix,iq) = D(iq,ix). */
void nko_setprec_local(double *dpc, const double *h1, const double *h2, int nx, int64_t nel,
                       const double *dt, const double *g1, const double *g2, const double *g3,
                       const double *g4, const double *g5, const double *g6, const double *bm1,
                       const int *ifdfrm)
{
    int n = nx * nx * nx;
#define DT(a, b) dt[(a) + nx * (b)]
#define IX(i, j, k) ((i) + nx * ((j) + nx * (k)))
    for (int64_t e = 0; e < nel; e++) {
        int64_t o = (int64_t)n * e;
        double *dp = dpc + o;
        for (int i = 0; i < n; i++) dp[i] = 0.0;
        for (int iq = 0; iq < nx; iq++) /* :415-421 */
            for (int iz = 0; iz < nx; iz++)
                for (int iy = 0; iy < nx; iy++)
                    for (int ix = 0; ix < nx; ix++)
                        dp[IX(ix, iy, iz)] = dp[IX(ix, iy, iz)] + g1[o + IX(iq, iy, iz)] * (DT(ix, iq) * DT(ix, iq));
        for (int iq = 0; iq < nx; iq++) /* :422-428 */
            for (int iz = 0; iz < nx; iz++)
                for (int iy = 0; iy < nx; iy++)
                    for (int ix = 0; ix < nx; ix++)
                        dp[IX(ix, iy, iz)] = dp[IX(ix, iy, iz)] + g2[o + IX(ix, iq, iz)] * (DT(iy, iq) * DT(iy, iq));
        for (int iq = 0; iq < nx; iq++) /* :430-436 */
            for (int iz = 0; iz < nx; iz++)
                for (int iy = 0; iy < nx; iy++)
                    for (int ix = 0; ix < nx; ix++)
                        dp[IX(ix, iy, iz)] = dp[IX(ix, iy, iz)] + g3[o + IX(ix, iy, iq)] * (DT(iz, iq) * DT(iz, iq));
        int dfrm = ifdfrm ? ifdfrm[e] : 1;
        if (dfrm) { /* :440-468 cross terms on the corners' faces */
            int L = nx - 1;
            for (int iy = 0; iy < nx; iy += L)
                for (int iz = 0; iz < nx; iz += L) {
                    dp[IX(0, iy, iz)] = dp[IX(0, iy, iz)] + g4[o + IX(0, iy, iz)] * DT(0, 0) * DT(iy, iy) +
                                        g5[o + IX(0, iy, iz)] * DT(0, 0) * DT(iz, iz);
                    dp[IX(L, iy, iz)] = dp[IX(L, iy, iz)] + g4[o + IX(L, iy, iz)] * DT(L, L) * DT(iy, iy) +
                                        g5[o + IX(L, iy, iz)] * DT(L, L) * DT(iz, iz);
                }
            for (int ix = 0; ix < nx; ix += L)
                for (int iz = 0; iz < nx; iz += L) {
                    dp[IX(ix, 0, iz)] = dp[IX(ix, 0, iz)] + g4[o + IX(ix, 0, iz)] * DT(0, 0) * DT(ix, ix) +
                                        g6[o + IX(ix, 0, iz)] * DT(0, 0) * DT(iz, iz);
                    dp[IX(ix, L, iz)] = dp[IX(ix, L, iz)] + g4[o + IX(ix, L, iz)] * DT(L, L) * DT(ix, ix) +
                                        g6[o + IX(ix, L, iz)] * DT(L, L) * DT(iz, iz);
                }
            for (int ix = 0; ix < nx; ix += L)
                for (int iy = 0; iy < nx; iy += L) {
                    dp[IX(ix, iy, 0)] = dp[IX(ix, iy, 0)] + g5[o + IX(ix, iy, 0)] * DT(0, 0) * DT(ix, ix) +
                                        g6[o + IX(ix, iy, 0)] * DT(0, 0) * DT(iy, iy);
                    dp[IX(ix, iy, L)] = dp[IX(ix, iy, L)] + g5[o + IX(ix, iy, L)] * DT(L, L) * DT(ix, ix) +
                                        g6[o + IX(ix, iy, L)] * DT(L, L) * DT(iy, iy);
                }
        }
        for (int i = 0; i < n; i++) { /* :491-492 col2(h1) ; addcol3(h2,bm1) */
            dp[i] = dp[i] * h1[o + i];
            dp[i] = dp[i] + h2[o + i] * bm1[o + i];
        }
    }
#undef DT
#undef IX
}

/* ------------------------------------------------------------------------- */
/* CG drivers                                                                 */
/* ------------------------------------------------------------------------- */

/* examples/bp5/bp5.usr:797-899 cggos with bpname='bp5' (dpc = 1, setprecn
 * :300-312), single rank (gop is the identity).  hist (may be NULL) receives
 * per iteration: pap, alpha, rtz (wv(2)), max|u-x1| (wv(1)) -> 4 doubles/iter.
 * Returns the number of iterations performed (maxit on exit of the loop, as
 * :889 iter=iter-1). */
int nko_cggos_bp5(double *u1, const double *rhs1, const double *x1, const double *rmult,
                  const double *v1mask, const int64_t *glo_num, const double *gf, int nx,
                  int64_t nel, const double *d, const double *dt, double tol, int maxit,
                  double *hist)
{
    int64_t n = (int64_t)nx * nx * nx * nel;
    double *dpc = malloc(sizeof(double) * 4 * (size_t)n);
    double *r1 = dpc + n, *p1 = dpc + 2 * n, *z1 = dpc + 3 * n, *ap1 = z1; /* equivalence :818-819 */
    for (int64_t i = 0; i < n; i++) dpc[i] = 1.0; /* setprecn */
    for (int64_t i = 0; i < n; i++) u1[i] = 0.0;
    for (int64_t i = 0; i < n; i++) r1[i] = rhs1[i];
    double wv1 = 0.0, wv2 = 0.0;
    for (int64_t i = 0; i < n; i++) { /* :839-843 */
        double s = rmult[i];
        p1[i] = dpc[i] * r1[i];
        wv1 = wv1 + s * p1[i] * r1[i];
    }
    double rpp1 = wv1, rpp2;
    int iter;
    for (iter = 1; iter <= maxit; iter++) {
        double pap = nko_axhm1_bp5(ap1, p1, gf, nx, nel, d, dt); /* :848 */
        nko_gs_op(ap1, glo_num, n, 1);                           /* :849 dssum */
        for (int64_t i = 0; i < n; i++) ap1[i] = ap1[i] * v1mask[i]; /* :850 xmask1 */
        double alph = rpp1 / pap;                                /* :853 */
        for (int64_t i = 0; i < n; i++) {                        /* :855-858 */
            u1[i] = u1[i] + alph * p1[i];
            r1[i] = r1[i] - alph * ap1[i];
        }
        wv1 = 0.0;
        wv2 = 0.0;
        for (int64_t i = 0; i < n; i++) { /* :861-866 */
            double s = fabs(u1[i] - x1[i]);
            wv1 = (wv1 > s) ? wv1 : s;
            z1[i] = dpc[i] * r1[i];
            wv2 = wv2 + rmult[i] * z1[i] * r1[i];
        }
        if (hist) {
            hist[4 * (iter - 1) + 0] = pap;
            hist[4 * (iter - 1) + 1] = alph;
            hist[4 * (iter - 1) + 2] = wv2;
            hist[4 * (iter - 1) + 3] = wv1;
        }
        double enorm = wv1;
        if (enorm < tol) { /* :870-874 */
            free(dpc);
            return iter;
        }
        rpp2 = rpp1;
        rpp1 = wv2;
        double beta1 = rpp1 / rpp2;
        for (int64_t i = 0; i < n; i++) p1[i] = z1[i] + beta1 * p1[i]; /* :882-884 */
    }
    free(dpc);
    return iter - 1; /* :889 */
}

/* core/hmholtz.f:611-846 cggo, Jacobi branch (kfldfdm<0, name != 'PRES'),
 * single rank, 3-D, param(18)=0, restol=0, param(22)>=0.  Performs setfast's
 * ifh2 test (:303-305), setprec (:690), the null-space correction (:705-720,
 * :747-749), and the iteration :726-816 with the exit rule of :778.
 * hist (may be NULL, length >= 3*niter) receives rtz1, rbn2, rho per iteration.
 * Returns niterhm. */
int nko_cggo(double *x, const double *f, const double *h1, const double *h2, const double *mask,
             const double *mult, const double *binv, const int64_t *glo_num, int nx, int64_t nel,
             const double *d, const double *dt, const double *g1, const double *g2,
             const double *g3, const double *g4, const double *g5, const double *g6,
             const double *bm1, const int *ifdfrm, double tin, int maxit, int istep, double *hist)
{
    const int maxcg = 900;
    int64_t n = (int64_t)nx * nx * nx * nel;
    double vol = 0.0;
    for (int64_t i = 0; i < n; i++) vol = vol + bm1[i]; /* volvm1 = glsum(bm1) core/coef.f */
    double tol = fabs(tin);
    int niter = maxit < maxcg ? maxit : maxcg;
    int ifh2 = 0;
    for (int64_t i = 0; i < n; i++)
        if (fabs(h2[i]) > 0.0) ifh2 = 1; /* setfast :303-305 */
    double *D = malloc(sizeof(double) * 5 * (size_t)n);
    double *r = D + n, *w = D + 2 * n, *p = D + 3 * n, *z = D + 4 * n;
    nko_setprec_local(D, h1, h2, nx, nel, dt, g1, g2, g3, g4, g5, g6, bm1, ifdfrm);
    nko_gs_op(D, glo_num, n, 1);
    for (int64_t i = 0; i < n; i++) D[i] = 1.0 / D[i]; /* invcol1 */
    for (int64_t i = 0; i < n; i++) {
        r[i] = f[i];
        x[i] = 0.0;
        p[i] = 0.0;
    }
    double fmax = 0.0;
    for (int64_t i = 0; i < n; i++)
        if (fabs(f[i]) > fmax) fmax = fabs(f[i]);
    if (fmax == 0.0) {
        free(D);
        return 0;
    }
    int ifmcor = 0;
    double h2max = h2[0], skmin = mask[0];
    for (int64_t i = 0; i < n; i++) {
        if (h2[i] > h2max) h2max = h2[i];
        if (mask[i] < skmin) skmin = mask[i];
    }
    if (skmin > 0 && h2max == 0) ifmcor = 1;
    double smean = 0.0, rmean;
    if (ifmcor) { /* :714-719 */
        double bsum = 0.0;
        for (int64_t i = 0; i < n; i++) bsum = bsum + bm1[i];
        smean = -1.0 / bsum;
        double s = 0.0;
        for (int64_t i = 0; i < n; i++) s = s + r[i] * mult[i]; /* glsc2 */
        rmean = smean * s;
        for (int64_t i = 0; i < n; i++) x[i] = bm1[i];
        nko_gs_op(x, glo_num, n, 1);
        for (int64_t i = 0; i < n; i++) r[i] = r[i] + rmean * x[i]; /* add2s2 */
        for (int64_t i = 0; i < n; i++) x[i] = 0.0;
    }
    double rtz1 = 1.0, rtz2, rho = 0.0, rbn2 = 0.0, rbn0 = 0.0;
    int iter;
    for (iter = 1; iter <= niter; iter++) {
        for (int64_t i = 0; i < n; i++) z[i] = r[i] * D[i]; /* :730 col3 */
        if (ifmcor) {                                      /* :747-749 */
            double s = 0.0;
            for (int64_t i = 0; i < n; i++) s = s + z[i] * bm1[i];
            rmean = smean * s;
            for (int64_t i = 0; i < n; i++) z[i] = z[i] + rmean;
        }
        rtz2 = rtz1;
        double s1 = 0.0, s2 = 0.0;
        for (int64_t i = 0; i < n; i++) s1 = s1 + z[i] * r[i] * mult[i];           /* vlsc3 navier4.f:323 */
        for (int64_t i = 0; i < n; i++) s2 = s2 + mult[i] * binv[i] * r[i] * r[i]; /* vlsc32 :848-856 */
        rtz1 = s1;
        rbn2 = sqrt(s2 / vol);
        if (iter == 1) rbn0 = rbn2;
        if (tin < 0) tol = fabs(tin) * rbn0; /* :765 */
        if (hist) {
            hist[3 * (iter - 1) + 0] = rtz1;
            hist[3 * (iter - 1) + 1] = rbn2;
            hist[3 * (iter - 1) + 2] = 0.0;
        }
        if (rbn2 <= tol && (iter > 1 || istep <= 5)) { /* :778 */
            niter = iter - 1;
            free(D);
            return niter;
        }
        double beta = rtz1 / rtz2;
        if (iter == 1) beta = 0.0;
        for (int64_t i = 0; i < n; i++) p[i] = beta * p[i] + z[i]; /* add2s1 :795 */
        nko_axhelm(w, p, h1, h2, ifh2, nx, nel, d, dt, g1, g2, g3, g4, g5, g6, bm1, ifdfrm);
        nko_gs_op(w, glo_num, n, 1);                         /* :797 */
        for (int64_t i = 0; i < n; i++) w[i] = w[i] * mask[i]; /* :798 */
        double s = 0.0;
        for (int64_t i = 0; i < n; i++) s = s + w[i] * p[i] * mult[i]; /* glsc3 :801 */
        rho = s;
        if (hist) hist[3 * (iter - 1) + 2] = rho;
        double alpha = rtz1 / rho, alphm = -alpha;
        for (int64_t i = 0; i < n; i++) x[i] = x[i] + alpha * p[i]; /* add2s2 :804 */
        for (int64_t i = 0; i < n; i++) r[i] = r[i] + alphm * w[i]; /* :805 */
    }
    niter = iter - 1; /* :817 */
    free(D);
    return niter;
}

/* examples/bp5/bp5.usr:422-447 glrdif */
double nko_glrdif(const double *x, const double *y, int64_t n)
{
    double dmx = 0, xmx = 0, ymx = 0;
    for (int64_t i = 0; i < n; i++) {
        double diff = fabs(x[i] - y[i]);
        dmx = dmx > diff ? dmx : diff;
        xmx = xmx > x[i] ? xmx : x[i];
        ymx = ymx > y[i] ? ymx : y[i];
    }
    xmx = xmx > ymx ? xmx : ymx;
    if (xmx > 0) return dmx / xmx;
    return -dmx;
}
