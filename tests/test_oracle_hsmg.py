"""Known-answer tests that pin the numpy oracle of the pressure preconditioner (oracle/hsmg.py).  The reference holds
no golden vectors for hsmg/gmres (parity unpinned); these checks are the analytic properties the
Fortran relies on."""
import numpy as np
import pytest

import oracle
from oracle import hsmg


def test_semhat_matches_speclib_and_is_a_stiffness_matrix():
    for n in (1, 3, 5, 7, 9):
        a, b, d, z = hsmg.semhat(n)
        zz, ww = oracle.zwgll(n + 1)
        assert np.array_equal(z, zz) and np.array_equal(b, ww)
        if n > 1:
            D, _ = oracle.dgll(zz)
            assert np.abs(d - D).max() <= 5e-13
        assert np.abs(a - a.T).max() <= 1e-13 and np.abs(a.sum(axis=1)).max() <= 1e-12   # symmetric, A 1 = 0
        x2 = z ** 2                                                                        # int (x^2)'^2 = 8/3 (exact: degree 2n-1 rule)
        if n >= 2:
            assert abs(x2 @ a @ x2 - 8.0 / 3.0) <= 1e-12


def test_interpolation_matrices_reproduce_polynomials():
    for lx1 in (4, 6, 8, 10):
        nx = hsmg.mg_orders(lx1)
        assert nx[0] == 1 and nx[-1] == lx1 - 1
        zs = [hsmg.semhat(n)[3] for n in nx]
        for l in range(len(nx) - 1):
            J = hsmg.intp_matrix(zs[l + 1], zs[l])
            for k in range(nx[l] + 1):
                assert np.abs(J @ zs[l] ** k - zs[l + 1] ** k).max() <= 1e-13
    assert hsmg.mg_orders(8) == [1, 3, 7] and hsmg.mg_orders(4) == [1, 3] and hsmg.mg_orders(6) == [1, 3, 5]


def test_fast1d_is_a_B_orthonormal_diagonalisation():
    a7, b7, _, _ = hsmg.semhat(7)
    for lbc, rbc in ((0, 0), (1, 2), (2, 1), (2, 2), (0, 1)):
        ll, lm, lr = 0.7, 1.0, 1.3
        A = hsmg.fast1d_a(lbc, rbc, ll, lm, lr, a7, 7)
        B = hsmg.fast1d_b(lbc, rbc, ll, lm, lr, b7, 7)
        lam, S = __import__("scipy.linalg").linalg.eigh(A, B, lower=False, driver="gv")
        assert np.abs(S.T @ B @ S - np.eye(10)).max() <= 1e-12
        assert np.abs(A @ S - B @ S * lam).max() <= 1e-10
        assert np.all(np.diff(lam) >= 0)


def test_fdm_inverts_the_dirichlet_problem_on_one_box_element():
    """One undeformed element with Dirichlet data on all six faces: the operator is separable, so the FDM solve is the
    exact inverse of the masked stiffness operator on the interior nodes."""
    case = oracle.Case(1, 1, 1, nx=8, dirichlet=(1, 1, 1, 1, 1, 1), hi=(1.0, 0.6, 1.7), rescale=False)
    mg = hsmg.H1MG(case, hsmg.box_fbc(case, (1, 1, 1, 1, 1, 1)))
    rng = np.random.default_rng(2)
    u = rng.standard_normal(case.n) * case.mask
    r = case.axhelm(u, np.ones(case.n), np.zeros(case.n)) * case.mask
    ext = np.zeros((1, 10, 10, 10))
    ext[:, 1:-1, 1:-1, 1:-1] = r.reshape(1, 8, 8, 8)
    e = mg.fdm_apply(ext, 2)[:, 1:-1, 1:-1, 1:-1].reshape(-1)
    assert np.abs(e - u).max() <= 1e-11 * np.abs(u).max()


def test_weights_masks_and_lengths_on_a_box():
    case = oracle.Case(3, 2, 2, nx=8, dirichlet=(0, 1, 0, 0, 0, 0), hi=(3.0, 1.0, 1.0), rescale=False)
    mg = hsmg.H1MG(case, hsmg.box_fbc(case, (2, 1, 2, 2, 2, 2)))
    assert np.allclose(mg.lm, [[1.0] * 12, [0.5] * 12, [0.5] * 12], rtol=1e-13)
    ex = np.arange(12) % 3
    assert np.allclose(mg.ll[0], np.where(ex > 0, 1.0, 0.0), atol=1e-13) and np.allclose(mg.lr[0], np.where(ex < 2, 1.0, 0.0), atol=1e-13)
    assert np.array_equal(mg.rstr_wt[2], case.mult)                     # restriction weight = inverse multiplicity
    assert np.array_equal(mg.mask[2], case.mask)                        # Dirichlet ('O') side zeroed at every level
    for l in (1, 2):
        cnt = 1.0 / mg.swt[l]
        assert np.abs(cnt - np.rint(cnt)).max() == 0 and cnt.min() >= 1   # overlap counts are small integers
        # interior of an element away from the overlap layers is counted once
        nh = mg.nh[l]
        if nh > 4:
            assert np.all(cnt.reshape(-1, nh, nh, nh)[:, 2:-2, 2:-2, 2:-2] == 1)


def test_coarse_operator_is_the_trilinear_stiffness_matrix():
    case = oracle.Case(2, 2, 2, nx=8, dirichlet=(1, 0, 0, 0, 0, 0))
    mg = hsmg.H1MG(case, hsmg.box_fbc(case, (1, 2, 2, 2, 2, 2)))
    A = mg.crs_A
    assert np.abs(A - A.T).max() <= 1e-13 and np.linalg.eigvalsh(A).min() > 0
    # element matrix of the trilinear Laplacian on a cube of side h: diagonal h/3, row sum 0
    h = 0.5
    assert np.allclose(np.diagonal(mg.crs_a, axis1=1, axis2=2), h / 3.0, rtol=1e-12)
    assert np.abs(mg.crs_a.sum(axis=2)).max() <= 1e-13
    b = np.random.default_rng(0).standard_normal(8 * case.nel) * mg.mask[0]
    x = mg.crs_solve(b)
    # x is continuous and solves Q^T A Q x = Q^T b on the unmasked dofs
    assert np.array_equal(mg.dssum(x, 0) * mg.rstr_wt[0], x) or np.allclose(mg.dssum(x, 0) * mg.rstr_wt[0], x, atol=1e-14)
    Ax = mg.dssum(np.einsum("eij,ej->ei", mg.crs_a, x.reshape(-1, 8)).reshape(-1), 0) * mg.mask[0]
    assert np.abs(Ax - mg.dssum(b, 0) * mg.mask[0]).max() <= 1e-11


@pytest.mark.parametrize("deform", [0.0, 0.03])
def test_preconditioned_gmres_beats_jacobi_pcg(deform):
    case = oracle.Case(3, 3, 2, nx=8, dirichlet=(0, 1, 0, 0, 0, 0), deform=deform)
    mg = hsmg.H1MG(case, hsmg.box_fbc(case, (2, 1, 2, 2, 2, 2)))
    rng = np.random.default_rng(0)
    n = case.n
    h1, h2 = np.ones(n), np.zeros(n)
    xe = case.dssum(rng.standard_normal(n)) * case.mult * case.mask
    b = case.dssum(case.axhelm(xe, h1, h2)) * case.mask
    x, it, hist, div0 = hsmg.hmh_gmres(case, mg, b, h1, h2, case.mask, case.mult, 1e-8, 100, history=True)
    assert it <= 30 and np.abs(x - xe).max() <= 1e-7 * np.abs(xe).max()
    assert np.all(np.diff(hist) <= 1e-14)                      # GMRES residuals are monotone
    _, itcg = case.cggo(b, h1, h2, tin=1e-8, maxit=900)
    assert itcg > 3 * it                                       # the multigrid preconditioner pays for itself
    # the preconditioner is a linear operator
    r1, r2 = rng.standard_normal(n), rng.standard_normal(n)
    z1, z2, z3 = mg.solve(r1.copy()), mg.solve(r2.copy()), mg.solve(2 * r1 - 3 * r2)
    assert np.abs(z3 - (2 * z1 - 3 * z2)).max() <= 1e-12 * np.abs(z3).max()


def test_all_neumann_null_space():
    case = oracle.Case(2, 2, 2, nx=6, dirichlet=(0, 0, 0, 0, 0, 0))
    mg = hsmg.H1MG(case, hsmg.box_fbc(case, (2, 2, 2, 2, 2, 2)), null_space=True)
    rng = np.random.default_rng(1)
    n = case.n
    h1, h2 = np.ones(n), np.zeros(n)
    xe = case.dssum(rng.standard_normal(n)) * case.mult
    b = case.dssum(case.axhelm(xe, h1, h2))
    x, it = hsmg.hmh_gmres(case, mg, b, h1, h2, case.mask, case.mult, 1e-9, 80, ifvcor=True)
    assert it < 40
    d = x - xe
    assert np.abs(d - d.mean()).max() <= 1e-6 * np.abs(xe).max()   # equal up to the constant null-space mode


def test_pnpn2_top_level_reduces_to_the_local_fdm_solve_on_one_element():
    """One element, no neighbours: local_solves_fdm is the interior of S D S^T applied to the zero-extended input and the
    overlap weights are 1; the whole hsmg_solve is linear."""
    case = oracle.Case(1, 1, 1, nx=8, dirichlet=(0, 1, 0, 0, 0, 0))
    fbc = hsmg.box_fbc(case, (2, 1, 2, 2, 2, 2))
    S, D = hsmg.standin_fastd(case, fbc)
    h = hsmg.Hsmg2(case, fbc, S, D)
    assert np.array_equal(h.owt, np.ones(216))
    rng = np.random.default_rng(0)
    v = rng.standard_normal(216)
    ext = np.zeros((1, 8, 8, 8))
    ext[:, 1:-1, 1:-1, 1:-1] = v.reshape(1, 6, 6, 6)
    t = np.einsum("eia,ejb,ekc,ekji->ecba", S[:, 0], S[:, 1], S[:, 2], ext) * D
    z = np.einsum("eai,ebj,eck,ekji->ecba", S[:, 0], S[:, 1], S[:, 2], t)[:, 1:-1, 1:-1, 1:-1].reshape(-1)
    assert np.abs(h.local_solves_fdm(v) - z).max() <= 1e-13 * np.abs(z).max()
    case = oracle.Case(3, 2, 2, nx=8, dirichlet=(0, 1, 0, 0, 0, 0), deform=0.02)
    fbc = hsmg.box_fbc(case, (2, 1, 2, 2, 2, 2))
    h = hsmg.Hsmg2(case, fbc, *hsmg.standin_fastd(case, fbc))
    cnt = 1.0 / h.owt
    assert np.array_equal(cnt, np.rint(cnt)) and cnt.max() == 4 and cnt.min() == 1
    r1, r2 = rng.standard_normal(216 * 12), rng.standard_normal(216 * 12)
    assert np.abs(h.solve(2 * r1 - 3 * r2) - (2 * h.solve(r1) - 3 * h.solve(r2))).max() <= 1e-10
