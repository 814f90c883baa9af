"""CPU oracle for the pressure preconditioner and its Krylov driver (TEST INFRASTRUCTURE ONLY).

numpy restatement of Nek5000's additive hybrid-Schwarz multigrid ``h1mg_solve`` (core/hsmg.f:1855-1949),
its setup ``h1mg_setup`` (core/hsmg.f:2234-2270), the single-level FDM pieces it is built from
(core/hsmg.f:368-553, 616-929, 1183-1319, 2216-2232, 3044-3100; core/fast3d.f:306-423, 802-877, 1215-1349,
1542-1617), the vertex-mesh coarse solve (core/navier8.f:83-233, 1648-1690; core/crs_xxt.c:926-965) and the
right-preconditioned GMRES ``hmh_gmres`` (core/gmres.f:284-545).

PARITY PINNED against the reference itself (see oracle/nek_oracle.c): tests/test_ref_pins.py compares h1mg_solve,
hmh_gmres, fdm_h1 / the Schwarz branch of cggo and hsmg_solve with the outputs of the reference's own Fortran
(oracle/_ref, transpiled by oracle/f77c.py) -- <= 1e-12 relative, identical iteration counts; the difference from
bit-exactness is the eigen-solver (the reference's 3rd_party/blasLapack dsygv vs SciPy's LAPACK dsygv,
scipy.linalg.eigh(driver="gv")) and the coarse factorisation.  tests/test_oracle_hsmg.py adds known-answer checks.

Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this module.

Array convention: flat Nek order u(i,j,k,e); ``_r(u, n)`` views it as [e, k, j, i].
"""
from __future__ import annotations

import numpy as np
import scipy.linalg

from . import lib, zwgll


def _r(u, n):
    return u.reshape(-1, n, n, n)


# ----------------------------------------------------------------------------- 1-D building blocks
def fd_weights_full(xx, x, m):
    """core/fast3d.f:1294-1349 (Fornberg).  Returns c[j, k], j = 0..n, k = 0..m."""
    n = len(x) - 1
    c = np.zeros((n + 1, m + 1))
    c1 = 1.0
    c4 = x[0] - xx
    c[0, 0] = 1.0
    for i in range(1, n + 1):
        mn = min(i, m)
        c2 = 1.0
        c5 = c4
        c4 = x[i] - xx
        for j in range(i):
            c3 = x[i] - x[j]
            c2 = c2 * c3
            if j == i - 1:
                for k in range(mn, 0, -1):
                    c[i, k] = c1 * (k * c[i - 1, k - 1] - c5 * c[i - 1, k]) / c2
                c[i, 0] = -c1 * c5 * c[i - 1, 0] / c2
            for k in range(mn, 0, -1):
                c[j, k] = (c4 * c[j, k] - k * c[j, k - 1]) / c3
            c[j, 0] = c4 * c[j, 0] / c3
        c1 = c2
    return c


def semhat(n):
    """core/fast3d.f:1215-1292 semhat, the parts h1mg uses: a (stiffness), b (mass = GLL weights), d, z."""
    z, b = zwgll(n + 1)
    d = np.zeros((n + 1, n + 1))
    for i in range(n + 1):
        d[i, :] = fd_weights_full(z[i], z, 1)[:, 1]
    a = np.zeros((n + 1, n + 1))
    for j in range(n + 1):
        for i in range(n + 1):
            s = 0.0
            for k in range(n + 1):
                s = s + d[k, i] * b[k] * d[k, j]
            a[i, j] = s
    return a, b, d, z


def intp_matrix(zf, zc):
    """core/hsmg.f:108-120 hsmg_setup_intpm: jh(nf, nc), Lagrange interpolation from zc to zf."""
    jh = np.zeros((len(zf), len(zc)))
    for i in range(len(zf)):
        jh[i, :] = fd_weights_full(zf[i], zc, 1)[:, 0]
    return jh


def mg_orders(lx1):
    """core/hsmg.f:2272-2337 h1mg_setup_mg_nx (3-D, lx2 = lx1: Pn-Pn)."""
    mgn2 = [1, 2, 2, 2, 2, 3, 3, 5, 5, 5]
    lmax = 2 if lx1 == 4 else 3
    mglx2 = 2 * (lx1 // 4) + 1
    if lx1 == 5:
        mglx2 = 3
    if lx1 <= 10:
        mglx2 = mgn2[min(lx1, 10) - 1]
    if lx1 == 8:
        mglx2 = 3
    mglx2 = min(3, mglx2)
    nx = [1, mglx2, mglx2 + 1]
    nx[lmax - 1] = lx1 - 1
    return nx[:lmax]


def fast1d_a(lbc, rbc, ll, lm, lr, ah, n):
    """core/hsmg.f:800-841 hsmg_setup_fast1d_a."""
    a = np.zeros((n + 3, n + 3))
    i0 = 1 if lbc == 1 else 0
    i1 = n - 1 if rbc == 1 else n
    fac = 2.0 / lm
    a[1, 1] = 1.0
    a[n + 1, n + 1] = 1.0
    for j in range(i0, i1 + 1):
        for i in range(i0, i1 + 1):
            a[i + 1, j + 1] = fac * ah[i, j]
    if lbc == 0:
        fac = 2.0 / ll
        a[0, 0] = fac * ah[n - 1, n - 1]
        a[1, 0] = fac * ah[n, n - 1]
        a[0, 1] = fac * ah[n - 1, n]
        a[1, 1] = a[1, 1] + fac * ah[n, n]
    else:
        a[0, 0] = 1.0
    if rbc == 0:
        fac = 2.0 / lr
        a[n + 1, n + 1] = a[n + 1, n + 1] + fac * ah[0, 0]
        a[n + 2, n + 1] = fac * ah[1, 0]
        a[n + 1, n + 2] = fac * ah[0, 1]
        a[n + 2, n + 2] = fac * ah[1, 1]
    else:
        a[n + 2, n + 2] = 1.0
    return a


def fast1d_b(lbc, rbc, ll, lm, lr, bh, n):
    """core/hsmg.f:843-879 hsmg_setup_fast1d_b."""
    b = np.zeros((n + 3, n + 3))
    i0 = 1 if lbc == 1 else 0
    i1 = n - 1 if rbc == 1 else n
    fac = 0.5 * lm
    b[1, 1] = 1.0
    b[n + 1, n + 1] = 1.0
    for i in range(i0, i1 + 1):
        b[i + 1, i + 1] = fac * bh[i]
    if lbc == 0:
        fac = 0.5 * ll
        b[0, 0] = fac * bh[n - 1]
        b[1, 1] = b[1, 1] + fac * bh[n]
    else:
        b[0, 0] = 1.0
    if rbc == 0:
        fac = 0.5 * lr
        b[n + 1, n + 1] = b[n + 1, n + 1] + fac * bh[0]
        b[n + 2, n + 2] = fac * bh[1]
    else:
        b[n + 2, n + 2] = 1.0
    return b


def fast1d(lbc, rbc, ll, lm, lr, ah, bh, n):
    """core/hsmg.f:775-798 hsmg_setup_fast1d: S (nl x nl, eigenvectors in columns, boundary rows zeroed), lam."""
    a = fast1d_a(lbc, rbc, ll, lm, lr, ah, n)
    b = fast1d_b(lbc, rbc, ll, lm, lr, bh, n)
    lam, s = scipy.linalg.eigh(a, b, lower=False, driver="gv")  # dsygv(1,'V','U'), core/hmholtz.f:1398
    nl = n + 3
    if lbc > 0:
        s[0, :] = 0.0
    if lbc == 1:
        s[1, :] = 0.0
    if rbc > 0:
        s[nl - 1, :] = 0.0
    if rbc == 1:
        s[nl - 2, :] = 0.0
    return s, lam


# ----------------------------------------------------------------------------- the preconditioner
class H1MG:
    """State of h1mg_setup for one oracle.Case.

    fbc[e, 0..5] = (lbr, rbr, lbs, rbs, lbt, rbt) of get_fast_bc (core/fast3d.f:802-877):
    0 interior/periodic ('E','P'), 1 Dirichlet for the pressure ('O','ON',...), 2 Neumann ('v','W','SYM',...).
    """

    def __init__(self, case, fbc, null_space=False):
        self.case = case
        self.E = case.nel
        self.lx1 = case.nx
        self.fbc = np.ascontiguousarray(fbc, dtype=np.int32).reshape(self.E, 6)
        self.null_space = bool(null_space)
        self.mg_nx = mg_orders(self.lx1)
        self.lmax = len(self.mg_nx)
        self.nh = [n + 1 for n in self.mg_nx]
        L = lib()
        E = self.E
        # h1mg_setup_semhat (core/hsmg.f:2339-2358)
        self.ah, self.bh, self.zh = [], [], []
        for n in self.mg_nx:
            a, b, d, z = semhat(n)
            self.ah.append(a), self.bh.append(b), self.zh.append(z)
        # hsmg_setup_intp (core/hsmg.f:83-106): jh[l] maps level l -> l+1 (0-based list index l)
        self.jh = [intp_matrix(self.zh[l + 1], self.zh[l]) for l in range(self.lmax - 1)]
        # h1mg_setup_dssum (core/hsmg.f:2360-2393): numbering of the nh^3 and (nh+2)^3 grids
        self.glo, self.glo_ext = [], []
        for nh in self.nh:
            g = np.zeros(nh ** 3 * E, dtype=np.int64)
            L.nko_setvert3d(g, nh, E, case.vertex, 1)
            self.glo.append(g)
            ge = np.zeros((nh + 2) ** 3 * E, dtype=np.int64)
            L.nko_setvert3d(ge, nh + 2, E, case.vertex, 1)
            self.glo_ext.append(ge)
        # h1mg_setup_wtmask -> hsmg_setup_rstr_wt (core/hsmg.f:987-1062): 1 / multiplicity on the element surface
        self.rstr_wt = []
        for l, nh in enumerate(self.nh):
            w = np.zeros((E, nh, nh, nh))
            w[:, 0], w[:, -1], w[:, :, 0], w[:, :, -1], w[:, :, :, 0], w[:, :, :, -1] = 1, 1, 1, 1, 1, 1
            w = self.dssum(w.reshape(-1), l)
            wt = np.ones_like(w)
            nzm = w != 0
            wt[nzm] = 1.0 / w[nzm]
            self.rstr_wt.append(wt)
        # mg_set_msk -> h1mg_setup_mask (core/hsmg.f:2395-2490): Dirichlet faces zeroed, gs multiply
        self.mask = []
        for l, nh in enumerate(self.nh):
            w = np.ones((E, nh, nh, nh))
            f = self.fbc
            w[f[:, 0] == 1, :, :, 0] = 0
            w[f[:, 1] == 1, :, :, -1] = 0
            w[f[:, 2] == 1, :, 0, :] = 0
            w[f[:, 3] == 1, :, -1, :] = 0
            w[f[:, 4] == 1, 0, :, :] = 0
            w[f[:, 5] == 1, -1, :, :] = 0
            self.mask.append(self.dssum(w.reshape(-1), l, op=2))
        self._lengths()
        # h1mg_setup_fdm (core/hsmg.f:632-664): levels 2..lmax
        self.fdm = [None] * self.lmax
        for l in range(1, self.lmax):
            self.fdm[l] = self._setup_fast(l)
        # h1mg_setup_schwarz_wt (core/hsmg.f:1217-1248, 3044-3100)
        self.swt = [None] * self.lmax
        for l in range(1, self.lmax):
            self.swt[l] = self._setup_schwarz_wt(l)
        self._setup_crs()

    # -- gather-scatter ------------------------------------------------------------------
    def dssum(self, u, l, op=1, ext=False):
        g = self.glo_ext[l] if ext else self.glo[l]
        u = np.ascontiguousarray(u, dtype=np.float64).copy()
        lib().nko_gs_op(u, g, len(u), op)
        return u

    # -- swap_lengths (core/fast3d.f:1542-1617) + plane_space (:306-423) -----------------
    def _lengths(self):
        c = self.case
        nx, E = self.lx1, self.E
        n2 = nx - 1
        x, y, z = _r(c.xm1, nx), _r(c.ym1, nx), _r(c.zm1, nx)
        w = c.w
        nin = nx - 2
        lm = np.zeros((3, E))
        for e in range(E):
            for d in range(3):
                s, ws = 0.0, 0.0
                for k in range(1, nin + 1):
                    for j in range(1, nin + 1):
                        wt = w[j - 1] * w[k - 1]  # the reference indexes wxm1 from 1 for point index 1 (0-based)
                        if d == 0:
                            a, b = (e, k, j, n2), (e, k, j, 0)
                        elif d == 1:
                            a, b = (e, k, n2, j), (e, k, 0, j)
                        else:
                            a, b = (e, n2, k, j), (e, 0, k, j)
                        s = s + wt / ((x[a] - x[b]) ** 2 + (y[a] - y[b]) ** 2 + (z[a] - z[b]) ** 2)
                        ws = ws + wt
                lm[d, e] = 1.0 / np.sqrt(s / ws)
        l = np.zeros((E, nx, nx, nx))
        for e in range(E):
            l[e, 1:n2, 1:n2, 0] = lm[0, e]
            l[e, 1:n2, 1:n2, n2] = lm[0, e]
            l[e, 1:n2, 0, 1:n2] = lm[1, e]
            l[e, 1:n2, n2, 1:n2] = lm[1, e]
            l[e, 0, 1:n2, 1:n2] = lm[2, e]
            l[e, n2, 1:n2, 1:n2] = lm[2, e]
        l = _r(self.dssum(l.reshape(-1), self.lmax - 1), nx)
        self.lm = lm
        self.ll = np.stack([l[:, 1, 1, 0] - lm[0], l[:, 1, 0, 1] - lm[1], l[:, 0, 1, 1] - lm[2]])
        self.lr = np.stack([l[:, 1, 1, n2] - lm[0], l[:, 1, n2, 1] - lm[1], l[:, n2, 1, 1] - lm[2]])

    # -- hsmg_setup_fast (core/hsmg.f:666-773) ---------------------------------------------
    def _setup_fast(self, l):
        n = self.mg_nx[l]
        nl = n + 3
        E = self.E
        S = np.zeros((E, 3, nl, nl))
        lam = np.zeros((E, 3, nl))
        D = np.zeros((E, nl, nl, nl))
        cache = {}
        for e in range(E):
            for d in range(3):
                key = (int(self.fbc[e, 2 * d]), int(self.fbc[e, 2 * d + 1]), self.ll[d, e], self.lm[d, e], self.lr[d, e])
                if key not in cache:
                    cache[key] = fast1d(key[0], key[1], key[2], key[3], key[4], self.ah[l], self.bh[l], n)
                S[e, d], lam[e, d] = cache[key]
            lr_, ls_, lt_ = lam[e, 0], lam[e, 1], lam[e, 2]
            eps = 1.0e-5 * (lr_[1:nl - 1].max() + ls_[1:nl - 1].max() + lt_[1:nl - 1].max())
            diag = lr_[None, None, :] + ls_[None, :, None] + lt_[:, None, None]
            D[e] = np.where(diag > eps, 1.0 / np.where(diag > eps, diag, 1.0), 0.0)
        return {"S": S, "lam": lam, "D": D, "nl": nl}

    def fdm_apply(self, r_ext, l):
        """core/hsmg.f:896-929 hsmg_do_fast (3-D): e = S D S^T r on (nh+2)^3 tiles."""
        f = self.fdm[l]
        S, D = f["S"], f["D"]
        # three successive 1-D contractions in the order of hsmg_tnsr3d_el (core/hsmg.f:309-324): r, then s, then t
        t = np.einsum("eia,ekji->ekja", S[:, 0], r_ext)
        t = np.einsum("ejb,ekja->ekba", S[:, 1], t)
        t = np.einsum("ekc,ekba->ecba", S[:, 2], t)
        t = D * t
        t = np.einsum("eai,ekji->ekja", S[:, 0], t)
        t = np.einsum("ebj,ekja->ekba", S[:, 1], t)
        return np.einsum("eck,ekba->ecba", S[:, 2], t)

    # -- hsmg_extrude (core/hsmg.f:368-423, 3-D) ------------------------------------------------
    @staticmethod
    def extrude(a1, l1, f1, a2, l2, f2):
        nx = a1.shape[-1]
        s = slice(1, nx - 1)
        a1[:, s, s, l1] = f1 * a1[:, s, s, l1] + f2 * a2[:, s, s, l2]
        a1[:, s, s, nx - 1 - l1] = f1 * a1[:, s, s, nx - 1 - l1] + f2 * a2[:, s, s, nx - 1 - l2]
        a1[:, s, l1, s] = f1 * a1[:, s, l1, s] + f2 * a2[:, s, l2, s]
        a1[:, s, nx - 1 - l1, s] = f1 * a1[:, s, nx - 1 - l1, s] + f2 * a2[:, s, nx - 1 - l2, s]
        a1[:, l1, s, s] = f1 * a1[:, l1, s, s] + f2 * a2[:, l2, s, s]
        a1[:, nx - 1 - l1, s, s] = f1 * a1[:, nx - 1 - l1, s, s] + f2 * a2[:, nx - 1 - l2, s, s]

    # -- h1mg_setup_schwarz_wt_1 (core/hsmg.f:3044-3100): full-array form of the weights ---------
    def _setup_schwarz_wt(self, l):
        nh, E = self.nh[l], self.E
        ne = nh + 2
        work = np.zeros((E, ne, ne, ne))
        ones = np.ones((E, ne, ne, ne))
        self.extrude(work, 0, 0.0, ones, 0, 1.0)
        ones = _r(self.dssum(ones.reshape(-1), l, ext=True), ne)
        self.extrude(ones, 0, 1.0, work, 0, -1.0)
        self.extrude(ones, 2, 1.0, ones, 0, 1.0)
        reg = np.ascontiguousarray(ones[:, 1:-1, 1:-1, 1:-1]).reshape(-1)
        reg = self.dssum(reg, l)
        # hsmg_schwarz_wt3d (core/hsmg.f:1285-1319) multiplies every node of the layers 1,2,n-1,n once by 1/count;
        # nodes outside those layers have count 1, so the weight is 1/count everywhere
        return 1.0 / reg

    # -- h1mg_schwarz (core/hsmg.f:425-494).  Masks r in place, as the reference does. -------------
    def schwarz(self, r, sigma, l):
        nh, E = self.nh[l], self.E
        ne = nh + 2
        r *= self.mask[l]
        work = np.zeros((E, ne, ne, ne))
        work[:, 1:-1, 1:-1, 1:-1] = _r(r, nh)
        self.extrude(work, 0, 0.0, work, 2, 1.0)
        work = _r(self.dssum(work.reshape(-1), l, ext=True), ne)
        self.extrude(work, 0, 1.0, work, 2, -1.0)
        e = self.fdm_apply(work, l)
        self.extrude(work, 0, 0.0, e, 0, 1.0)
        e = _r(self.dssum(e.reshape(-1), l, ext=True), ne)
        self.extrude(e, 0, 1.0, work, 0, -1.0)
        self.extrude(e, 2, 1.0, e, 0, 1.0)
        out = np.ascontiguousarray(e[:, 1:-1, 1:-1, 1:-1]).reshape(-1)
        out = self.dssum(out, l)
        out *= self.mask[l]
        out *= self.swt[l]
        out *= sigma
        return out

    # -- h1mg_rstr / hsmg_intp (core/hsmg.f:2216-2232, 205-212) ---------------------------------------
    def rstr(self, r, l, ifdssum):
        """r (level l+1) -> level l: J^T (rstr_wt * r) [, dssum].  l is the 0-based coarse level."""
        J = self.jh[l]
        v = _r(r * self.rstr_wt[l + 1], self.nh[l + 1])
        out = np.einsum("ia,jb,kc,ekji->ecba", J, J, J, v).reshape(-1)
        if ifdssum:
            out = self.dssum(out, l)
        return np.ascontiguousarray(out)

    def intp(self, uc, l):
        """level l -> l+1 (0-based coarse level l)."""
        J = self.jh[l]
        return np.ascontiguousarray(np.einsum("ai,bj,ck,ekji->ecba", J, J, J, _r(uc, self.nh[l])).reshape(-1))

    # -- coarse grid (core/navier8.f:83-233, 1648-1690; core/crs_xxt.c:926-965) --------------------
    def _setup_crs(self):
        c = self.case
        E, nx = self.E, self.lx1
        z0, z1 = 0.5 * (1 - c.z), 0.5 * (1 + c.z)
        basis = []
        for j in range(1, 9):  # gen_crs_basis, navier8.f:1692-1733
            zr = z1 if j % 2 == 0 else z0
            zs = z1 if j in (3, 4, 7, 8) else z0
            zt = z1 if j > 4 else z0
            basis.append(np.einsum("k,j,i->kji", zt, zs, zr).reshape(-1))
        self.crs_basis = np.array(basis)
        ones, zeros = np.ones(c.n), np.zeros(c.n)
        a = np.zeros((E, 8, 8))
        for j in range(8):
            w2 = c.axhelm(np.tile(basis[j], E), ones, zeros).reshape(E, -1)
            for i in range(8):
                a[:, i, j] = w2 @ basis[i]
        self.crs_a = a
        ids = self.glo[0].copy()
        ids[self.mask[0] == 0] = 0  # set_jl_crs_mask
        uniq = np.unique(ids[ids != 0])
        self.crs_ids = ids
        self.crs_dof = np.searchsorted(uniq, ids)  # valid where ids != 0
        n = len(uniq)
        A = np.zeros((n, n))
        dof = self.crs_dof.reshape(E, 8)
        live = (ids != 0).reshape(E, 8)
        for e in range(E):
            for i in range(8):
                if not live[e, i]:
                    continue
                for j in range(8):
                    if live[e, j]:
                        A[dof[e, i], dof[e, j]] += a[e, i, j]
        self.crs_A = A
        self.crs_n = n

    def crs_solve(self, b):
        ids = self.crs_ids
        live = ids != 0
        rhs = np.zeros(self.crs_n)
        np.add.at(rhs, self.crs_dof[live], b[live])
        if self.null_space:  # pin one dof, then remove the mean over the distinct dofs (crs_xxt.c:944-955)
            x = np.zeros(self.crs_n)
            x[:-1] = np.linalg.solve(self.crs_A[:-1, :-1], rhs[:-1])
            x -= x.mean()
        else:
            x = np.linalg.solve(self.crs_A, rhs)
        out = np.zeros_like(b)
        out[live] = x[self.crs_dof[live]]
        return out

    # -- h1mg_solve (core/hsmg.f:1855-1949), additive (if_hybrid = .false., core/gmres.f:330) -------
    def solve(self, rhs):
        """Returns z; rhs is masked in place (h1mg_schwarz_part1 masks its input)."""
        c = self.case
        sigma = 1.0
        lm = self.lmax - 1
        z = self.schwarz(rhs, sigma, lm)
        r = rhs.copy()
        e = [None] * self.lmax
        for l in range(lm - 1, 0, -1):
            r = self.rstr(r, l, True)
            e[l] = self.schwarz(r, sigma, l)
        r = self.rstr(r, 0, False)
        r *= self.mask[0]
        e[0] = self.crs_solve(r)
        e[0] *= self.mask[0]
        for l in range(1, lm):
            e[l] = e[l] + self.intp(e[l - 1], l - 1)
        z = z + self.intp(e[lm - 1], lm - 1)
        return c.dssum(z) * c.mult  # dsavg, core/ic.f:1871


def box_fbc(case, bc):
    """fbc[e, 6] for a box case: bc = six codes (x-, x+, y-, y+, z-, z+) applied on the box sides (1 Dirichlet for the
    pressure, 2 Neumann); interior and periodic faces are 0."""
    E = case.nel
    f = np.zeros((E, 6), dtype=np.int32)
    ex = np.arange(E) % case.nelx
    ey = (np.arange(E) // case.nelx) % case.nely
    ez = np.arange(E) // (case.nelx * case.nely)
    for d, (idx, nd) in enumerate(((ex, case.nelx), (ey, case.nely), (ez, case.nelz))):
        if bc[2 * d] != 0:
            f[idx == 0, 2 * d] = bc[2 * d]
        if bc[2 * d + 1] != 0:
            f[idx == nd - 1, 2 * d + 1] = bc[2 * d + 1]
    return f


# ----------------------------------------------------------------------------- hmh_gmres
def hmh_gmres(case, mg, res, h1, h2, pmask, wt, tol, maxit, m=30, ifvcor=False, history=False):
    """core/gmres.f:304-545 with ml = mu = 1 (uzawa_gmres_split, :252-271), if_hyb = .false.; `tol` is tolpss.
    Returns (x, iterations[, rnorm history])."""
    n = case.n
    norm_fac = 1.0 / np.sqrt(case.bm1().sum())
    glsc3 = lambda a, b: float(np.sum(a * b * wt))

    def ortho(x):
        if ifvcor:
            x -= x.sum() / n
        return x

    def ax(x):
        return case.dssum(case.axhelm(x, h1, h2)) * pmask

    x = np.zeros(n)
    V = np.zeros((m + 1, n))
    Z = np.zeros((m, n))
    H = np.zeros((m + 1, m))
    cg, sg, gam = np.zeros(m), np.zeros(m), np.zeros(m + 1)
    it, conv, hist = 0, False, []
    div0 = 0.0
    while not conv:
        if it == 0:
            r = res.copy()
        else:
            r = res - ax(x)
        gam[0] = np.sqrt(glsc3(r, r))
        if it == 0:
            div0 = gam[0] * norm_fac
        if gam[0] == 0.0:
            break
        V[0] = r / gam[0]
        j = 0
        for j in range(m):
            it += 1
            w = V[j].copy()
            Z[j] = ortho(mg.solve(w))
            w = ax(Z[j])
            for i in range(j + 1):
                H[i, j] = glsc3(w, V[i])
            for i in range(j + 1):
                w = w - H[i, j] * V[i]
            for i in range(j):
                t = H[i, j]
                H[i, j] = cg[i] * t + sg[i] * H[i + 1, j]
                H[i + 1, j] = -sg[i] * t + cg[i] * H[i + 1, j]
            alpha = np.sqrt(glsc3(w, w))
            if alpha == 0.0:
                conv = True
                break
            l = np.sqrt(H[j, j] * H[j, j] + alpha * alpha)
            t = 1.0 / l
            cg[j] = H[j, j] * t
            sg[j] = alpha * t
            H[j, j] = l
            gam[j + 1] = -sg[j] * gam[j]
            gam[j] = cg[j] * gam[j]
            rnorm = abs(gam[j + 1]) * norm_fac
            hist.append(rnorm)
            if it + 1 > maxit or rnorm < tol:
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
    x = ortho(x)
    if history:
        return x, it, np.array(hist), div0
    return x, it


def hmh_flex_cg(case, mg, res, h1, h2, pmask, wt, tol, maxit, ifvcor=False):
    """core/hmholtz.f:2164-2290: flexible PCG with h1mg_solve (param(42) = 2); `tol` is tolpss.  Returns (x, iterations)."""
    n = case.n
    vol = case.bm1().sum()
    glsc3 = lambda a, b: float(np.sum(a * wt * b))
    ax = lambda x: case.dssum(case.axhelm(x, h1, h2)) * pmask
    r, r1, p, x = res.copy(), np.zeros(n), np.zeros(n), np.zeros(n)
    rho1, it = 1.0, 0
    for _ in range(maxit):
        z = mg.solve(r)                       # masks r in place, as the reference does
        r1 = r1 - r
        rho0 = rho1
        rho1 = glsc3(z, r)
        rho2 = -glsc3(z, r1)
        beta = rho2 / rho0
        r1 = r.copy()
        p = beta * p + z
        w = ax(p)
        alpha = rho1 / glsc3(w, p)
        x = x + alpha * p
        r = r - alpha * w
        rnorm = np.sqrt(float(np.sum(r * r * wt)) / vol)
        it += 1
        if rnorm < tol:
            break
    if ifvcor:
        x = x - x.sum() / n
    return x, it


# ----------------------------------------------------------------------------- fdm_h1 (single-level Schwarz / FDM)
class FdmH1:
    """core/hmholtz.f:1028-1112 set_fdm_prec_h1A_gen, :1114-1220 set_fdm_prec_h1A_els, :1222-1290 set_fdm_prec_h1b and
    :937-1026 fdm_h1 for one field.

    face_internal[e, 6]: 1 where cbc is 'E  ', 'P  ' or 'p  ' (faces in symmetric order r-, r+, s-, s+, t-, t+);
    mask: the field's Dirichlet mask (decides Dirichlet vs Neumann on the remaining faces)."""

    def __init__(self, case, face_internal, mask):
        n = case.nx
        self.case, self.n = case, n
        z, w = case.z, case.w
        D = case.D
        delta = abs(z[1] - z[0])
        bbh, aah = 0.5 * delta, 1.0 / delta
        self.dd = np.zeros((9, n))
        self.fds = np.zeros((9, n, n))
        l = 0
        for right in (1, 2, 3):
            for left in (1, 2, 3):
                bb = np.diag(w).astype(np.float64)
                aa = D.T @ (bb @ D)
                if left == 1:
                    bb[0, 0] += bbh
                    aa[0, 0] += aah
                elif left == 2:
                    bb[0, 0] = 1.0
                    aa[:, 0] = 0.0
                    aa[0, :] = 0.0
                    aa[0, 0] = 1.0
                if right == 1:
                    bb[-1, -1] += bbh
                    aa[-1, -1] += aah
                elif right == 2:
                    bb[-1, -1] = 1.0
                    aa[:, -1] = 0.0
                    aa[-1, :] = 0.0
                    aa[-1, -1] = 1.0
                lam, s = scipy.linalg.eigh(aa, bb, lower=False, driver="gv")
                self.dd[l], self.fds[l] = lam, s
                l += 1
        E = case.nel
        fi = np.asarray(face_internal).reshape(E, 6)
        m = _r(np.asarray(mask), n)
        self.ktype = np.zeros((E, 3), dtype=np.int32)
        pts = (((1, 1, 0), (1, 1, n - 1)), ((1, 0, 1), (1, n - 1, 1)), ((0, 1, 1), (n - 1, 1, 1)))  # [k, j, i] of k1, k2
        for e in range(E):
            for d in range(3):
                code = []
                for side in (0, 1):
                    k, j, i = pts[d][side]
                    if fi[e, 2 * d + side]:
                        code.append(1)
                    elif m[e, k, j, i] == 0:
                        code.append(2)
                    else:
                        code.append(3)
                self.ktype[e, d] = code[0] + 3 * (code[1] - 1)
        x, y, zz = _r(case.xm1, n), _r(case.ym1, n), _r(case.zm1, n)
        self.elsize = np.zeros((3, E))
        w3 = [w[0] * w[:, None] * w[None, :]] * 3  # wxm1(i)*wxm1(j)*wxm1(k) with the normal index fixed at 1
        for e in range(E):
            for d in range(3):
                if d == 0:
                    a, b = (e, slice(None), slice(None), n - 1), (e, slice(None), slice(None), 0)
                elif d == 1:
                    a, b = (e, slice(None), n - 1, slice(None)), (e, slice(None), 0, slice(None))
                else:
                    a, b = (e, n - 1, slice(None), slice(None)), (e, 0, slice(None), slice(None))
                dl2 = (x[a] - x[b]) ** 2 + (y[a] - y[b]) ** 2 + (zz[a] - zz[b]) ** 2
                self.elsize[d, e] = np.sqrt((dl2 * w3[d]).sum() / w3[d].sum()) / 2.0

    def set_prec_h1b(self, h1, h2):
        n, E = self.n, self.case.nel
        d = np.zeros((E, n, n, n))
        h1b = h1.reshape(E, -1).sum(axis=1) / n ** 3
        h2b = h2.reshape(E, -1).sum(axis=1) / n ** 3
        for e in range(E):
            k1, k2, k3 = self.ktype[e] - 1
            s = self.elsize[:, e]
            vol = s[0] * s[1] * s[2]
            vl1, vl2, vl3 = s[1] * s[2] / s[0], s[0] * s[2] / s[1], s[0] * s[1] / s[2]
            den = h1b[e] * (vl1 * self.dd[k1][None, None, :] + vl2 * self.dd[k2][None, :, None] + vl3 * self.dd[k3][:, None, None]) + h2b[e] * vol
            nzm = den != 0
            d[e][nzm] = 1.0 / den[nzm]
        return d.reshape(-1)

    def apply(self, r, d, mask):
        n, E = self.n, self.case.nel
        S = self.fds[self.ktype - 1]  # [E, 3, n, n]
        t = np.einsum("eia,ejb,ekc,ekji->ecba", S[:, 0], S[:, 1], S[:, 2], _r(r, n))
        t = t * _r(d, n)
        z = np.einsum("eai,ebj,eck,ekji->ecba", S[:, 0], S[:, 1], S[:, 2], t).reshape(-1)
        return self.case.dssum(z) * mask


def crs_solve_h1(case, mg, v, mult):
    """core/navier8.f:1490-1535 crs_solve_h1: uf = J crs_solve(J^T (vmult * vf)) with the bilinear vertex basis
    h1_basis(i, 0:2) = (1 -+ z_i)/2 (:1537-1550), restriction :1605-1646, prolongation :1555-1603 (r, then s, then t)."""
    nx, E = case.nx, case.nel
    J = np.stack([0.5 * (1.0 - case.z), 0.5 * (1.0 + case.z)], axis=1)        # [lx1, 2]
    uf = (v * mult).reshape(E, nx, nx, nx)                                     # [e, k, j, i]
    t = np.einsum("ia,ekji->ekja", J, uf)                                      # mxm(h1_basist, 2, uf, lx1, v, ...)
    t = np.einsum("jb,ekja->ekba", J, t)
    vc = np.einsum("kc,ekba->ecba", J, t).reshape(-1)
    uc = mg.crs_solve(vc).reshape(E, 2, 2, 2)
    t = np.einsum("ia,ecba->ecbi", J, uc)
    t = np.einsum("jb,ecbi->ecji", J, t)
    return np.einsum("kc,ecji->ekji", J, t).reshape(-1)


def cggo_schwarz(case, fdm, f, h1, h2, mask, tin, maxit, istep=1, history=False, pres_mg=None, ifvcor=False):
    """core/hmholtz.f:611-846 cggo with kfldfdm >= 0 (Schwarz branch :737-745) without the null-space correction.  With
    pres_mg (an H1MG holding the coarse solver): the name = 'PRES' form of param(42) = 1 -- z = fdm_h1(r) + crs_solve_h1(r)
    (:741-744), then ortho(z) (:747-748, core/navier1.f:223-256: the plain mean over all entries goes when ifvcor).
    Returns (x, niter[, hist rows (rtz1, rbn2, rho)])."""
    n = case.n
    mult, binv = case.mult, case.binv()
    vol = case.bm1().sum()
    tol = abs(tin)
    niter = min(maxit, 900)
    d = fdm.set_prec_h1b(h1, h2)
    r, x, p = f.copy(), np.zeros(n), np.zeros(n)
    hist = []
    if np.abs(f).max() == 0.0:
        return (x, 0, np.zeros((0, 3))) if history else (x, 0)
    rtz1, rho, rbn0 = 1.0, 0.0, 0.0
    it_done = niter
    for it in range(1, niter + 1):
        z = fdm.apply(r, d, mask)
        if pres_mg is not None:
            z = z + crs_solve_h1(case, pres_mg, r, mult)
            if ifvcor:
                z = z - z.sum() / n
        rtz2 = rtz1
        rtz1 = float(np.sum(z * r * mult))
        rbn2 = float(np.sqrt(np.sum(mult * binv * r * r) / vol))
        if it == 1:
            rbn0 = rbn2
        if tin < 0:
            tol = abs(tin) * rbn0
        row = [rtz1, rbn2, 0.0]
        hist.append(row)
        if rbn2 <= tol and (it > 1 or istep <= 5):
            it_done = it - 1
            break
        beta = 0.0 if it == 1 else rtz1 / rtz2
        p = z + beta * p
        w = case.dssum(case.axhelm(p, h1, h2)) * mask
        rho = float(np.sum(w * p * mult))
        row[2] = rho
        alpha = rtz1 / rho
        x = x + alpha * p
        r = r - alpha * w
    if history:
        return x, it_done, np.array(hist)
    return x, it_done


# ----------------------------------------------------------------------------- hsmg_solve (Pn-Pn-2)
class Hsmg2:
    """core/hsmg.f:22-47 hsmg_setup and :1376-1602 hsmg_solve (additive) with the top level of core/fasts.f:2-94
    (local_solves_fdm, fastdm1, dface_ext, dface_add1si, s_face_to_int, init_weight_op, do_weight_op) on the lx2 = lx1-2
    Gauss grid.  The top-level FDM data S[e,3,lx1,lx1] (eigenvectors in columns), D[e,lx1,lx1,lx1] is an INPUT, as it is for
    the library (common /fastd/ produced by gen_fast)."""

    def __init__(self, case, fbc, S, D, null_space=False):
        self.case = case
        self.low = H1MG(case, fbc, null_space)      # levels below the top are built by the same routines
        lx1 = case.nx
        assert hsmg_orders_pnpn2(lx1)[:-1] == mg_orders(lx1)[:-1]
        self.lmax = self.low.lmax
        self.n2 = lx1 - 2
        self.S, self.D = S, D
        zgl = np.polynomial.legendre.leggauss(self.n2)[0]
        self.jtop = intp_matrix(zgl, self.low.zh[self.lmax - 2])   # hsmg_setup_intp with mg_zh(lmax) = zglhat (hsmg.f:57)
        # init_weight_op (fasts.f:310-413)
        E, n = case.nel, lx1
        l = np.zeros((E, n, n, n))
        l[:, 1:-1, 1:-1, 1], l[:, 1:-1, 1:-1, n - 2] = 1, 1
        l[:, 1:-1, 1, 1:-1], l[:, 1:-1, n - 2, 1:-1] = 1, 1
        l[:, 1, 1:-1, 1:-1], l[:, n - 2, 1:-1, 1:-1] = 1, 1
        self.dface_ext(l)
        l = _r(case.dssum(l.reshape(-1)), n)
        self.dface_add1si(l, -1.0)
        self.s_face_to_int(l, 1.0)
        cnt = l[:, 1:-1, 1:-1, 1:-1].copy()
        w = np.ones_like(cnt)
        outer = np.zeros_like(cnt, dtype=bool)
        outer[:, :, :, 0] = outer[:, :, :, -1] = outer[:, :, 0, :] = outer[:, :, -1, :] = outer[:, 0, :, :] = outer[:, -1, :, :] = True
        w[outer] = 1.0 / cnt[outer]
        self.owt = w.reshape(-1)

    @staticmethod
    def dface_ext(x):
        s = slice(1, -1)
        x[:, s, 0, s], x[:, s, -1, s] = x[:, s, 1, s], x[:, s, -2, s]
        x[:, s, s, 0], x[:, s, s, -1] = x[:, s, s, 1], x[:, s, s, -2]
        x[:, 0, s, s], x[:, -1, s, s] = x[:, 1, s, s], x[:, -2, s, s]

    @staticmethod
    def dface_add1si(x, c):
        s = slice(1, -1)
        x[:, s, 0, s] += c * x[:, s, 1, s]
        x[:, s, -1, s] += c * x[:, s, -2, s]
        x[:, s, s, 0] += c * x[:, s, s, 1]
        x[:, s, s, -1] += c * x[:, s, s, -2]
        x[:, 0, s, s] += c * x[:, 1, s, s]
        x[:, -1, s, s] += c * x[:, -2, s, s]

    @staticmethod
    def s_face_to_int(x, c):
        s = slice(1, -1)
        x[:, s, 1, s] = c * x[:, s, 0, s] + x[:, s, 1, s]
        x[:, s, -2, s] = c * x[:, s, -1, s] + x[:, s, -2, s]
        x[:, s, s, 1] = c * x[:, s, s, 0] + x[:, s, s, 1]
        x[:, s, s, -2] = c * x[:, s, s, -1] + x[:, s, s, -2]
        x[:, 1, s, s] = c * x[:, 0, s, s] + x[:, 1, s, s]
        x[:, -2, s, s] = c * x[:, -1, s, s] + x[:, -2, s, s]

    def local_solves_fdm(self, v):
        c, n, n2 = self.case, self.case.nx, self.n2
        v1 = np.zeros((c.nel, n, n, n))
        v1[:, 1:-1, 1:-1, 1:-1] = _r(v, n2)
        self.dface_ext(v1)
        v1 = _r(c.dssum(v1.reshape(-1)), n)
        self.dface_add1si(v1, -1.0)
        S = self.S
        t = np.einsum("eia,ejb,ekc,ekji->ecba", S[:, 0], S[:, 1], S[:, 2], v1) * self.D
        v1 = np.einsum("eai,ebj,eck,ekji->ecba", S[:, 0], S[:, 1], S[:, 2], t)
        self.s_face_to_int(v1, -1.0)
        v1 = _r(c.dssum(v1.reshape(-1)), n)
        self.s_face_to_int(v1, 1.0)
        return np.ascontiguousarray(v1[:, 1:-1, 1:-1, 1:-1]).reshape(-1) * self.owt

    def solve(self, r):
        lo, lm = self.low, self.lmax - 1
        e = self.local_solves_fdm(r)
        w = r.copy()
        el = [None] * self.lmax
        for l in range(lm - 1, 0, -1):
            if l + 1 == lm:   # hsmg_rstr from the top: no weights (hsmg.f:220-222), J^T, dssum
                J = self.jtop
                rl = lo.dssum(np.ascontiguousarray(np.einsum("ia,jb,kc,ekji->ecba", J, J, J, _r(w, self.n2)).reshape(-1)), l)
            else:
                rl = lo.rstr(w, l, True)
            el[l] = lo.schwarz(rl.copy(), 1.0, l)    # hsmg_schwarz + hsmg_schwarz_wt; masks only its copy
            w = rl
        if lm == 1:
            J = self.jtop
            r1 = np.ascontiguousarray(np.einsum("ia,jb,kc,ekji->ecba", J, J, J, _r(w, self.n2)).reshape(-1))
        else:
            r1 = lo.rstr(w, 0, False)
        r1 = r1 * lo.mask[0]
        el[0] = lo.crs_solve(r1) * lo.mask[0]
        for l in range(1, lm):
            el[l] = el[l] + lo.intp(el[l - 1], l - 1)
        J = self.jtop
        e = e + np.einsum("ai,bj,ck,ekji->ecba", J, J, J, _r(el[lm - 1], lo.nh[lm - 1])).reshape(-1)
        if lo.null_space:
            e = e - e.sum() / len(e)
        return e


def semhat_weighted(n):
    """core/fast3d.f:1181-1213 load_semhat_weighted: bh (GLL weights), and jgl, dgl (velocity GLL nodes -> pressure GL
    nodes interpolation / derivative, :1280-1289) pre-multiplied by the GL weights bgl.  Returns bh[n+1], jgl[n-1,n+1],
    dgl[n-1,n+1]."""
    z, bh = zwgll(n + 1)
    zgl, bgl = np.polynomial.legendre.leggauss(n - 1)
    jgl, dgl = np.zeros((n - 1, n + 1)), np.zeros((n - 1, n + 1))
    for i in range(n - 1):
        w = fd_weights_full(zgl[i], z, 1)
        jgl[i], dgl[i] = w[:, 0], w[:, 1]
    return bh, bgl[:, None] * jgl, bgl[:, None] * dgl


def fast1d_sem_op(b0, b1, l, r, ll, lm, lr, bh, jgl, jscl):
    """core/fast3d.f:1410-1540 set_up_fast_1D_sem_op: G = J B^-1 J^T restricted to an element plus one node either side.
    jgl[i-1, k] is the Fortran jgl(i,k), i = 1..n-1, k = 0..n."""
    n = len(bh) - 1
    J = lambda i, k: jgl[i - 1, k]
    gl, gm, gr = (1.0, 1.0, 1.0) if jscl == 0 else (0.5 * ll, 0.5 * lm, 0.5 * lr)
    gll, glm, gmm, gmr, grr = gl * gl, gl * gm, gm * gm, gm * gr, gr * gr
    bm, bl, br = np.zeros(n + 1), np.zeros(n + 1), np.zeros(n + 1)
    for i in range(1, n):
        bm[i] = 2.0 / (lm * bh[i])
    if b0 == 0:
        bm[0] = 0.5 * lm * bh[0]
        if l:
            bm[0] = bm[0] + 0.5 * ll * bh[n]
        bm[0] = 1.0 / bm[0]
    if b1 == n:
        bm[n] = 0.5 * lm * bh[n]
        if r:
            bm[n] = bm[n] + 0.5 * lr * bh[0]
        bm[n] = 1.0 / bm[n]
    if l:
        for i in range(n):
            bl[i] = 2.0 / (ll * bh[i])
        bl[n] = bm[0]
    if r:
        for i in range(1, n + 1):
            br[i] = 2.0 / (lr * bh[i])
        br[0] = bm[n]
    g = np.zeros((n + 1, n + 1))
    for j in range(1, n):
        for i in range(1, n):
            for k in range(b0, b1 + 1):
                g[i, j] = g[i, j] + gmm * J(i, k) * bm[k] * J(j, k)
    if l:
        for i in range(1, n):
            g[i, 0] = glm * J(i, 0) * bm[0] * J(n - 1, n)
            g[0, i] = g[i, 0]
        for i in range(n + 1):
            g[0, 0] = g[0, 0] + gll * J(n - 1, i) * bl[i] * J(n - 1, i)
    else:
        g[0, 0] = 1.0
    if r:
        for i in range(1, n):
            g[i, n] = gmr * J(i, n) * bm[n] * J(1, 0)
            g[n, i] = g[i, n]
        for i in range(n + 1):
            g[n, n] = g[n, n] + grr * J(1, i) * br[i] * J(1, i)
    else:
        g[n, n] = 1.0
    return g


def fast1d_sem(lbc, rbc, ll, lm, lr, bh, jgl, dgl):
    """core/fast3d.f:1351-1408 set_up_fast_1D_sem: eigen-system of the 1-D E~ x = lam B~ x; codes of get_fast_bc with
    bsym = 3 (0 element, 1 outflow, 2 wall, 3 symmetry).  Returns S (eigenvectors in columns, boundary rows zeroed), lam."""
    n = len(bh) - 1
    eb0 = 1 if lbc in (2, 3) else 0
    eb1 = n - 1 if rbc in (2, 3) else n
    bb0 = 1 if lbc == 2 else 0
    bb1 = n - 1 if rbc == 2 else n
    l, r = lbc == 0, rbc == 0
    e = fast1d_sem_op(eb0, eb1, l, r, ll, lm, lr, bh, dgl, 0)
    b = fast1d_sem_op(bb0, bb1, l, r, ll, lm, lr, bh, jgl, 1)
    lam, s = scipy.linalg.eigh(e, b, lower=False, driver="gv")     # generalev -> dsygv(1,'V','U')
    if not l:
        s[0, :] = 0.0
    if not r:
        s[n, :] = 0.0
    return s, lam


def gen_fast(case, fbc):
    """core/fast3d.f:2-140 gen_fast with param(44) = 0 (the default: top-level Schwarz on restrictions of E): the
    common /fastd/ data of the Pn-Pn-2 preconditioner.  Returns S[e,3,lx1,lx1], D[e,lx1,lx1,lx1] (D = df)."""
    fbc = np.asarray(fbc).reshape(case.nel, 6)
    mg = H1MG(case, np.where(fbc == 3, 2, fbc))   # swap_lengths: ll, lm, lr (symmetry faces count as walls there)
    nl = case.nx
    bh, jgl, dgl = semhat_weighted(nl - 1)
    E = case.nel
    S, D = np.zeros((E, 3, nl, nl)), np.zeros((E, nl, nl, nl))
    cache = {}
    for e in range(E):
        lam = []
        for d in range(3):
            key = (int(fbc[e, 2 * d]), int(fbc[e, 2 * d + 1]), mg.ll[d, e], mg.lm[d, e], mg.lr[d, e])
            if key not in cache:
                cache[key] = fast1d_sem(*key, bh, jgl, dgl)
            S[e, d], l1 = cache[key]
            lam.append(l1)
        eps = 1e-5 * (lam[0][1:-1].max() + lam[1][1:-1].max() + lam[2][1:-1].max())
        diag = lam[0][None, None, :] + lam[1][None, :, None] + lam[2][:, None, None]
        D[e] = np.where(diag > eps, 1.0 / np.where(diag > eps, diag, 1.0), 0.0)
    return S, D


def hsmg_orders_pnpn2(lx1):
    """core/hsmg.f:1604-1664 hsmg_setup_mg_nx."""
    mgn2 = [1, 2, 2, 2, 2, 3, 3, 5, 5, 5]
    lmax = 2 if lx1 == 4 else 3
    mglx2 = 2 * ((lx1 - 2) // 4) + 1
    if lx1 == 5:
        mglx2 = 3
    if lx1 <= 10:
        mglx2 = mgn2[min(lx1, 10) - 1]
    if lx1 == 8:
        mglx2 = 3
    nx = [1, mglx2, mglx2 + 1]
    nx[lmax - 1] = lx1 - 1
    return nx[:lmax]


def standin_fastd(case, fbc):
    """A realistic stand-in for common /fastd/ (gen_fast is not restated): the 1-D systems of hsmg_setup_fast1d for the
    order lx1-3 (so that nl = lx1), with the element lengths of swap_lengths.  Returns S[e,3,lx1,lx1], D[e,lx1,lx1,lx1]."""
    mg = H1MG(case, fbc)
    n = case.nx - 3
    a, b, _, _ = semhat(n)
    E, nl = case.nel, case.nx
    S, D = np.zeros((E, 3, nl, nl)), np.zeros((E, nl, nl, nl))
    for e in range(E):
        lam = []
        for d in range(3):
            s, l = fast1d(int(mg.fbc[e, 2 * d]), int(mg.fbc[e, 2 * d + 1]), mg.ll[d, e], mg.lm[d, e], mg.lr[d, e], a, b, n)
            S[e, d] = s
            lam.append(l)
        diag = lam[0][None, None, :] + lam[1][None, :, None] + lam[2][:, None, None]
        eps = 1e-5 * (lam[0][1:-1].max() + lam[1][1:-1].max() + lam[2][1:-1].max())
        D[e] = np.where(diag > eps, 1.0 / np.where(diag > eps, diag, 1.0), 0.0)
    return S, D
