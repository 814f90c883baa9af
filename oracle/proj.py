"""CPU oracle of the residual projection around the Helmholtz / pressure solves (TEST INFRASTRUCTURE ONLY).

numpy restatement of core/navier4.f: hsolve (:562-634), hmhzpf (:513-560), project1 (:636-733), project1_a (:735-793),
iproj_chk (:795-826), proj_matvec (:828-845), proj_ortho (:848-945), proj_ortho_full_cgs2 (:1060-1125), project2 /
project2_a (:1129-1199), proj_get_ivar (:1201-1236), givens_rotation / hypot (:1308-1344), and chktcg1
(core/hmholtz.f:527-609).  Pinned against the reference's own output by tests/test_ref_pins.py (hsolve, hsolve_pres cases).
"""
from __future__ import annotations

import numpy as np


def givens_rotation(a, b):
    if b != 0.0:
        c, d = abs(a), abs(b)
        x = max(c, d)
        t = (1.0 / x) * min(c, d)
        h = x * np.sqrt(1.0 + t * t)
        dd = 1.0 / h
        return abs(a) * dd, np.copysign(dd, a) * b, np.copysign(1.0, a) * h
    return 1.0, 0.0, a


def chktcg1(case, tol, res, h1, h2, mask, mult, binv, vol):
    """core/hmholtz.f:527-609 (double precision: eps = 1e-13; acondno = 10 since eigaa = 0)."""
    eps, acondno = 1.0e-13, 10.0
    rinit = np.sqrt(np.sum(binv * res * res * mult) / vol)
    rmin = eps * rinit
    if tol < rmin:
        tol = rmin
    one = np.ones(case.n)
    bctest = abs(np.sum(one * mask * mult) - np.sum(one * one * mult))
    w2 = case.axhelm(one, h1, h2)
    bcrob = np.sqrt(np.sum(w2 * w2 * case.bm1()) / vol)
    if bctest < 0.1 and bcrob < eps * acondno:
        tolmin = rinit * eps * 10.0
        if tol < tolmin:
            tol = tolmin
    return tol


class Projection:
    """State of one solver name: X, B = A X (columns), xbar, bbar, h1old, h2old, m (ivar(2)), mmx (ivar(1))."""

    def __init__(self, case, mask, w, mxprev=20):
        self.case, self.mask, self.w = case, mask, w
        self.mmx = (mxprev - 4) // 2
        n = case.n
        self.X, self.B = np.zeros((self.mmx, n)), np.zeros((self.mmx, n))
        self.xbar, self.bbar = np.zeros(n), np.zeros(n)
        self.h1old, self.h2old = np.zeros(n), np.zeros(n)
        self.m = 0

    def dot(self, a, b):
        return float(np.sum(a * self.w * b))

    def matvec(self, x, h1, h2):
        return self.case.dssum(self.case.axhelm(x, h1, h2)) * self.mask

    def sym(self, j, k):
        return 0.5 * (self.dot(self.X[j], self.B[k]) + self.dot(self.B[j], self.X[k]))

    def ortho_full_cgs2(self):
        m, tol = self.m, 1.0e-7
        if m <= 0:
            return
        flag = [0] * m
        for _ in range(2):
            for k in range(m - 1, -1, -1):
                alpha = {j: self.sym(j, k) for j in range(m - 1, k - 1, -1)}
                for j in range(m - 1, k, -1):
                    self.X[k] -= alpha[j] * self.X[j]
                    self.B[k] -= alpha[j] * self.B[j]
                normp = np.sqrt(alpha[k])
                normk = np.sqrt(self.dot(self.X[k], self.B[k]))
                if normk > tol * normp:
                    self.X[k] *= 1.0 / normk
                    self.B[k] *= 1.0 / normk
                    flag[k] = 1
                else:
                    flag[k] = 0
        k = 0
        for j in range(m):
            if flag[j]:
                if k < j:
                    self.X[k], self.B[k] = self.X[j].copy(), self.B[j].copy()
                k += 1
        self.m = k

    def ortho(self):
        m = self.m
        if m <= 0:
            return
        last = m - 1
        alpha = np.array([self.sym(k, last) for k in range(m)])
        nrm = np.sqrt(alpha[last])
        for k in range(m - 1):
            self.X[last] -= alpha[k] * self.X[k]
            self.B[last] -= alpha[k] * self.B[k]
        beta = np.array([self.sym(k, last) for k in range(m - 1)])
        for k in range(m - 1):
            self.X[last] -= beta[k] * self.X[k]
            self.B[last] -= beta[k] * self.B[k]
            alpha[k] += beta[k]
        alpha[last] = np.sqrt(self.dot(self.X[last], self.B[last]))
        if alpha[last] > 1.0e-7 * nrm:
            s1 = 1.0 / alpha[last]
            self.X[last] *= s1
            self.B[last] *= s1
            for k in range(m - 1, 0, -1):
                h = k - 1
                c, s, nrm = givens_rotation(alpha[h], alpha[k])
                alpha[h] = nrm
                xh, xk = self.X[h].copy(), self.X[k].copy()
                self.X[h], self.X[k] = c * xh + s * xk, -s * xh + c * xk
                bh, bk = self.B[h].copy(), self.B[k].copy()
                self.B[h], self.B[k] = c * bh + s * bk, -s * bh + c * bk
        else:
            self.m = m - 1

    def project1(self, b, h1, h2):
        """b is modified in place."""
        if self.m <= 0:
            return
        dh = max(np.abs(h1 - self.h1old).max(), np.abs(h2 - self.h2old).max())
        if dh > 0:
            self.h1old, self.h2old = h1.copy(), h2.copy()
            for j in range(self.m):
                self.B[j] = self.matvec(self.X[j], h1, h2)
            self.ortho_full_cgs2()
            if self.m <= 0:
                return
        m = self.m
        for rnd in range(2):
            alpha = [self.dot(self.X[k], b) for k in range(m)]
            for k in range(m):
                if rnd == 0 and k == 0:
                    self.xbar, self.bbar = alpha[0] * self.X[0], alpha[0] * self.B[0]
                else:
                    self.xbar = self.xbar + alpha[k] * self.X[k]
                    self.bbar = self.bbar + alpha[k] * self.B[k]
                b -= alpha[k] * self.B[k]

    def project2(self, x, h1, h2):
        """x is modified in place."""
        if self.m > 0:
            x += self.xbar
        self.m = min(self.m + 1, self.mmx)
        self.X[self.m - 1] = x
        self.B[self.m - 1] = self.matvec(x, h1, h2)
        self.ortho()


def hsolve_projected(case, P, r, h1, h2, tol, maxit, istep, binv, vol, solver):
    """The projection branch of hsolve (navier4.f:603-627): returns (u, r_out, niter).  solver(rhs, tol) -> (x, niter) is the
    cggo call of hmhzpf (Jacobi PCG, or hmh_gmres for 'PRES')."""
    r = r * P.mask
    r = case.dssum(r)
    P.project1(r, h1, h2)
    t = chktcg1(case, tol, r, h1, h2, P.mask, P.w, binv, vol)      # hmhzpf, param(22) = 0
    u, it = solver(r, t)
    P.project2(u, h1, h2)
    return u, r, it
