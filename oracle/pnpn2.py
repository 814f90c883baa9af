"""CPU oracle of the Pn-Pn-2 pressure operator (TEST INFRASTRUCTURE ONLY).

numpy restatement of core/navier1.f: cdtp (:330-536) / opgradt (:4095-4114), multd (:538-714) / opdiv (:4064-4093),
opbinv (:775-850), cdabdtp (:258-293), chktcg2 (:1089-1160) and of uzawa_gmres (core/gmres.f:2-237) for the 3-D,
non-axisymmetric, ifsplit = .false. branch.  Arrays are flat Nek order; mesh 2 is the (lx1-2)^3 Gauss grid.
Pinned against the reference's own output by tests/test_ref_pins.py (case "eop").
"""
from __future__ import annotations

import numpy as np


class Mesh2:
    """ixm12, dxm12: [a, i] = ixm12(a,i); w3m2 (lx2^3 flat); met: rxm2, sxm2, txm2, rym2, ..., tzm2 (flat, E*lx2^3)."""

    def __init__(self, case, ixm12, dxm12, w3m2, met, bm2=None, bm2inv=None, volvm2=None):
        self.case = case
        self.I, self.D = np.asarray(ixm12), np.asarray(dxm12)
        self.n2, self.n1 = self.I.shape
        self.E = case.nel
        sh = (self.E, self.n2, self.n2, self.n2)
        self.w3 = np.asarray(w3m2).reshape(self.n2, self.n2, self.n2)
        self.met = [np.asarray(m).reshape(sh) for m in met]
        self.bm2 = None if bm2 is None else np.asarray(bm2).reshape(-1)
        self.bm2inv = None if bm2inv is None else np.asarray(bm2inv).reshape(-1)
        self.volvm2 = volvm2

    def _up(self, f, Ax, Ay, Az):      # f[e,c,b,a] -> [e,k,j,i]
        return np.einsum("ai,bj,ck,ecba->ekji", Ax, Ay, Az, f, optimize=True)

    def _down(self, u, Ax, Ay, Az):    # u[e,k,j,i] -> [e,c,b,a]
        return np.einsum("ai,bj,ck,ekji->ecba", Ax, Ay, Az, u, optimize=True)

    def cdtp(self, x, isd):
        wx = self.w3[None] * x.reshape(self.E, self.n2, self.n2, self.n2)
        r, s, t = self.met[3 * isd:3 * isd + 3]
        I, D = self.I, self.D
        out = self._up(wx * r, D, I, I)
        out = out + self._up(wx * s, I, D, I)
        out = out + self._up(wx * t, I, I, D)
        return out.reshape(-1)

    def multd(self, u, isd):
        u = u.reshape(self.E, self.n1, self.n1, self.n1)
        r, s, t = self.met[3 * isd:3 * isd + 3]
        I, D = self.I, self.D
        dx = self._down(u, D, I, I) * r
        dx = dx + self._down(u, I, D, I) * s
        dx = dx + self._down(u, I, I, D) * t
        return (dx * self.w3[None]).reshape(-1)

    def opgradt(self, p):
        return [self.cdtp(p, isd) for isd in range(3)]

    def opdiv(self, u):
        out = self.multd(u[0], 0)
        out = out + self.multd(u[1], 1)
        return out + self.multd(u[2], 2)

    def opbinv(self, inp, h2inv, masks):
        """Returns (out[3], inp_after[3])."""
        c = self.case
        a = [c.dssum(inp[k] * masks[k]) for k in range(3)]
        d = c.dssum(c.bm1() / h2inv)
        return [a[k] * (1.0 / d) for k in range(3)], a

    def cdabdtp(self, wp, h2inv, masks):
        """intype = 1."""
        tb, _ = self.opbinv(self.opgradt(wp), h2inv, masks)
        return self.opdiv(tb)

    def cdabdtp_helm(self, wp, h1, h2, masks, tolhs, nmxv, istep=20):
        """intype = 0 / -1: the three velocity solves of ophinv (hmholtz -> cggo) between D^T and D."""
        c = self.case
        ta = self.opgradt(wp)
        tb = []
        for k in range(3):
            rhs = c.dssum(ta[k]) * masks[k]
            x, _ = c.cggo(rhs, h1, h2, mask=masks[k], tin=tolhs, maxit=nmxv, istep=istep)
            tb.append(x)
        return self.opdiv(tb)


def chktcg2(M, tol, res, ifvcor=False, prelax=0.0, tolpdf=0.0):
    """core/navier1.f:1089-1154."""
    eps = prelax if prelax != 0.0 else 1.0e-10
    ta = res * M.bm2inv
    rinit = np.sqrt(np.sum(ta * ta * M.bm2) / M.volvm2)
    if rinit < tol:
        return tol
    rmin = tolpdf if tolpdf > 0.0 else eps * rinit
    if tol < rmin:
        tol = rmin
    if ifvcor:
        tolmin = abs(np.sum(res)) * 100.0
        if tol < tolmin:
            tol = tolmin
    return tol


def uzawa_gmres(M, precond, res, h2inv, masks, tolps, param21=0.0, istep=1, m=30, ifvcor=False, ntotg=None):
    """core/gmres.f:2-237 with intype = 1; precond(w) = hsmg_solve.  Returns (x, iter)."""
    n2 = len(res)
    ml, mu = np.sqrt(M.bm2inv), np.sqrt(M.bm2)
    norm_fac = 1.0 / np.sqrt(M.volvm2)
    tolps = chktcg2(M, tolps, res, ifvcor)
    if param21 > 0 and tolps > abs(param21):
        tolps = abs(param21)
    if istep == 0:
        tolps = 1.0e-4
    tolpss = tolps
    E = lambda p: M.cdabdtp(p, h2inv, masks)
    x = np.zeros(n2)
    V, Z = np.zeros((m + 1, n2)), np.zeros((m, n2))
    H = np.zeros((m + 1, m))
    cg, sg, gam = np.zeros(m), np.zeros(m), np.zeros(m + 1)
    it, conv = 0, False
    while not conv and it < 100:
        r = ml * res if it == 0 else ml * (res - E(x))
        gam[0] = np.sqrt(np.sum(r * r))
        if it == 0:
            div0 = gam[0] * norm_fac
            if param21 < 0:
                tolpss = abs(param21) * div0
        if gam[0] == 0.0:
            break
        V[0] = r / gam[0]
        j = 0
        for j in range(m):
            it += 1
            Z[j] = precond(mu * V[j])
            w = ml * E(Z[j])
            for i in range(j + 1):
                H[i, j] = np.sum(w * V[i])
            for i in range(j + 1):
                w = w - H[i, j] * V[i]
            for i in range(j):
                t = H[i, j]
                H[i, j] = cg[i] * t + sg[i] * H[i + 1, j]
                H[i + 1, j] = -sg[i] * t + cg[i] * H[i + 1, j]
            alpha = np.sqrt(np.sum(w * w))
            if alpha == 0.0:
                conv = True
                break
            l = np.sqrt(H[j, j] * H[j, j] + alpha * alpha)
            t = 1.0 / l
            cg[j], sg[j] = H[j, j] * t, alpha * t
            H[j, j] = l
            gam[j + 1] = -sg[j] * gam[j]
            gam[j] = cg[j] * gam[j]
            rnorm = abs(gam[j + 1]) * norm_fac
            if rnorm < tolpss:
                conv = True
                break
            if j == m - 1:
                break
            V[j + 1] = w / alpha
        k = j + 1
        cvec = np.zeros(k)
        for q in range(k - 1, -1, -1):
            t = gam[q]
            for i in range(k - 1, q, -1):
                t = t - H[q, i] * cvec[i]
            cvec[q] = t / H[q, q]
        for i in range(k):
            x = x + cvec[i] * Z[i]
    if ifvcor:
        x = x - np.sum(x) / ntotg
    return x, it
