/*
 * oracle/bp5_cpu.c -- "restated CPU baseline, N threads" for bench.py.
 *
 * TEST / MEASUREMENT INFRASTRUCTURE ONLY (see oracle/nek_oracle.c header).
 * The reference's MPI ranks become OpenMP threads: elements are block-partitioned
 * over the threads exactly as Nek block-partitions them over ranks, each thread
 * runs the reference's element loop (examples/bp5/bp5.usr:1309-1341 axhm1_bp5 ->
 * :1278-1307 ax_e_bp5), the gather-scatter is a shared-memory gs over
 * precomputed id groups (gslib gs_op semantics, core/dssum.f:79), and the
 * reductions are OpenMP reductions in place of gop (core/comm_mpi.f:216-259).
 * Same iteration (bp5.usr:847-885), same fixed iteration count, timed like
 * bp5.usr:366-370.  Built -O3 -march=native -fopenmp as BASELINE.md section 4
 * prescribes; this is a reported baseline, not an optimisation target.
 */
#include <math.h>
#include <omp.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

/* ax_e_bp5 for nx=8 written so gcc can vectorise; same operation order as
 * mxf8 (core/mxm_std.f:174-191) up to FMA contraction. */
static void ax_e(double *restrict w, const double *restrict u, const double *restrict g,
                 const double *restrict d, const double *restrict dt, int nx, double *restrict wk)
{
    int n = nx * nx * nx, nxy = nx * nx;
    double *ur = wk, *us = wk + n, *ut = wk + 2 * n;
    /* ur = D u : (nx x nx)(nx x nx^2) */
    for (int j = 0; j < nxy; j++)
        for (int i = 0; i < nx; i++) {
            double s = 0.0;
            for (int k = 0; k < nx; k++) s += d[i + nx * k] * u[k + nx * j];
            ur[i + nx * j] = s;
        }
    for (int kz = 0; kz < nx; kz++)
        for (int j = 0; j < nx; j++)
            for (int i = 0; i < nx; i++) {
                double s = 0.0;
                for (int k = 0; k < nx; k++) s += u[i + nx * k + nxy * kz] * dt[k + nx * j];
                us[i + nx * j + nxy * kz] = s;
            }
    for (int j = 0; j < nx; j++)
        for (int i = 0; i < nxy; i++) {
            double s = 0.0;
            for (int k = 0; k < nx; k++) s += u[i + nxy * k] * dt[k + nx * j];
            ut[i + nxy * j] = s;
        }
    for (int i = 0; i < n; i++) {
        double wr = g[6 * i + 0] * ur[i] + g[6 * i + 1] * us[i] + g[6 * i + 2] * ut[i];
        double ws = g[6 * i + 1] * ur[i] + g[6 * i + 3] * us[i] + g[6 * i + 4] * ut[i];
        double wt = g[6 * i + 2] * ur[i] + g[6 * i + 4] * us[i] + g[6 * i + 5] * ut[i];
        ur[i] = wr;
        us[i] = ws;
        ut[i] = wt;
    }
    for (int j = 0; j < nxy; j++)
        for (int i = 0; i < nx; i++) {
            double s = 0.0;
            for (int k = 0; k < nx; k++) s += dt[i + nx * k] * ur[k + nx * j];
            w[i + nx * j] = s;
        }
    for (int kz = 0; kz < nx; kz++)
        for (int j = 0; j < nx; j++)
            for (int i = 0; i < nx; i++) {
                double s = w[i + nx * j + nxy * kz];
                for (int k = 0; k < nx; k++) s += us[i + nx * k + nxy * kz] * d[k + nx * j];
                w[i + nx * j + nxy * kz] = s;
            }
    for (int j = 0; j < nx; j++)
        for (int i = 0; i < nxy; i++) {
            double s = w[i + nxy * j];
            for (int k = 0; k < nx; k++) s += ut[i + nxy * k] * d[k + nx * j];
            w[i + nxy * j] = s;
        }
}

/* Group CSR built by the caller (sorted by id): off[ngrp+1], idx[off[ngrp]]. */
static void gs_add(double *u, const int64_t *off, const int32_t *idx, int64_t ngrp)
{
#pragma omp parallel for schedule(static)
    for (int64_t g = 0; g < ngrp; g++) {
        double s = 0.0;
        for (int64_t q = off[g]; q < off[g + 1]; q++) s += u[idx[q]];
        for (int64_t q = off[g]; q < off[g + 1]; q++) u[idx[q]] = s;
    }
}

/* Runs `maxit` fixed iterations of bp5.usr cggos (dpc=1, tol<0) and returns the
 * wall-clock seconds of the loop.  u1 receives the iterate. */
double nkb_cpu_cggos(double *u1, const double *rhs1, const double *rmult, const double *mask,
                     const int64_t *off, const int32_t *idx, int64_t ngrp, const double *gf, int nx,
                     int64_t nel, const double *d, const double *dt, int maxit, int nthreads)
{
    int64_t nxyz = (int64_t)nx * nx * nx, n = nxyz * nel;
    if (nthreads > 0) omp_set_num_threads(nthreads);
    double *r1 = malloc(sizeof(double) * (size_t)n);
    double *p1 = malloc(sizeof(double) * (size_t)n);
    double *ap = malloc(sizeof(double) * (size_t)n);
    double rpp1 = 0.0;
#pragma omp parallel for reduction(+ : rpp1) schedule(static)
    for (int64_t i = 0; i < n; i++) {
        u1[i] = 0.0;
        r1[i] = rhs1[i];
        p1[i] = r1[i];
        rpp1 += rmult[i] * p1[i] * r1[i];
    }
    struct timespec t0, t1;
    clock_gettime(CLOCK_MONOTONIC, &t0);
    for (int iter = 1; iter <= maxit; iter++) {
        double pap = 0.0;
#pragma omp parallel reduction(+ : pap)
        {
            double *wk = malloc(sizeof(double) * 3 * (size_t)nxyz);
#pragma omp for schedule(static)
            for (int64_t e = 0; e < nel; e++) {
                ax_e(ap + nxyz * e, p1 + nxyz * e, gf + 6 * nxyz * e, d, dt, nx, wk);
                double s = 0.0;
                for (int i = 0; i < nxyz; i++) s += p1[nxyz * e + i] * ap[nxyz * e + i];
                pap += s;
            }
            free(wk);
        }
        gs_add(ap, off, idx, ngrp);
        double alph = rpp1 / pap, rz = 0.0;
#pragma omp parallel for reduction(+ : rz) schedule(static)
        for (int64_t i = 0; i < n; i++) {
            double a = ap[i] * mask[i];
            u1[i] = u1[i] + alph * p1[i];
            r1[i] = r1[i] - alph * a;
            rz += rmult[i] * r1[i] * r1[i];
        }
        double beta = rz / rpp1;
        rpp1 = rz;
#pragma omp parallel for schedule(static)
        for (int64_t i = 0; i < n; i++) p1[i] = r1[i] + beta * p1[i];
    }
    clock_gettime(CLOCK_MONOTONIC, &t1);
    free(r1);
    free(p1);
    free(ap);
    return (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec);
}

int nkb_cpu_max_threads(void) { return omp_get_max_threads(); }
