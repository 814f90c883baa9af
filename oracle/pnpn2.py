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
