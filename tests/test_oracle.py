"""Pins the CPU oracle (oracle/nek_oracle.c) against the reference's own mesh fixtures
(tests/golden/bp5_fixture.npz <- examples/bp5/bp5.{re2,ma2}) and against analytic known-answer
tests.  The reference holds no golden vectors for axhelm/dssum/cggo: parity unpinned."""
import os

import numpy as np
import pytest

import oracle


def classes(ids):
    """Canonical equivalence-class labelling of an id vector (0 = unshared stays 0)."""
    out = np.zeros(len(ids), dtype=np.int64)
    nz = np.nonzero(ids)[0]
    _, first, inv = np.unique(ids[nz], return_index=True, return_inverse=True)
    out[nz] = nz[first][inv] + 1  # label = 1 + smallest local index in the class
    return out


# --------------------------------------------------------------------------- speclib
def test_gll_points_weights_known_answers():
    # Known closed forms: N=2 -> (-1,0,1),(1/3,4/3,1/3); N=3 -> +-1/sqrt5, (1/6,5/6)
    z, w = oracle.zwgll(3)
    assert np.allclose(z, [-1, 0, 1], atol=1e-15) and np.allclose(w, [1 / 3, 4 / 3, 1 / 3], atol=1e-15)
    z, w = oracle.zwgll(4)
    assert np.allclose(z, [-1, -1 / np.sqrt(5), 1 / np.sqrt(5), 1], atol=1e-15)
    assert np.allclose(w, [1 / 6, 5 / 6, 5 / 6, 1 / 6], atol=1e-15)
    for nx in range(2, 17):
        z, w = oracle.zwgll(nx)
        n = nx - 1
        # interior GLL points are the roots of P_n'
        c = np.zeros(n + 1); c[n] = 1
        roots = np.sort(np.polynomial.legendre.Legendre(c).deriv().roots())
        assert np.allclose(z[1:-1], roots, atol=1e-13)
        pn = np.polynomial.legendre.Legendre(c)(z)
        assert np.allclose(w, 2.0 / (n * (n + 1) * pn ** 2), rtol=1e-13)
        # exact for degree 2n-1
        for k in range(0, 2 * n):
            assert abs((w * z ** k).sum() - (0 if k % 2 else 2.0 / (k + 1))) < 1e-13


def test_dgll_differentiates_polynomials_exactly():
    for nx in (2, 4, 8, 10, 12):
        z, _ = oracle.zwgll(nx)
        D, Dt = oracle.dgll(z)
        assert np.array_equal(D.T, Dt)
        for k in range(nx):
            exact = k * z ** max(k - 1, 0) if k else np.zeros(nx)
            assert np.allclose(D @ z ** k, exact, atol=2e-12 * nx * nx)


def test_mxm_matches_numpy_and_order():
    rng = np.random.default_rng(1)
    a, b = rng.standard_normal((8, 8)), rng.standard_normal((8, 64))
    c = oracle.mxm(a, b)
    assert np.allclose(c, a @ b, rtol=0, atol=1e-13)
    # left-to-right order: exact reproduction with an explicit python loop
    ref = a[:, [0]] * b[[0], :]
    for k in range(1, 8):
        ref = ref + a[:, [k]] * b[[k], :]
    assert np.array_equal(c, ref)


def test_ran1_numerical_recipes_sequence():
    # NR 2nd ed. ran1 with idum=-1: 0.4159994, 0.0919649, 0.7564105, 0.5297002, 0.9304365
    x = np.zeros(5)
    oracle.lib().nko_rand_fld(x, 5)
    assert np.allclose(x, [0.4159994, 0.0919649, 0.7564105, 0.5297002, 0.9304365], atol=5e-8)


# --------------------------------------------------------------------------- fixtures from the reference
@pytest.fixture(scope="module")
def fixture(golden_dir):
    return np.load(os.path.join(golden_dir, "bp5_fixture.npz"))


def test_box_mesh_matches_bp5_re2(fixture):
    c = oracle.Case(10, 10, 10, nx=3, rescale=False)
    for mine, ref in ((c.xc, fixture["xc"]), (c.yc, fixture["yc"]), (c.zc, fixture["zc"])):
        # genbox accumulates dx (0.30000000000000004 style): identical element order and corner order,
        # coordinates equal to rounding
        assert np.abs(mine.reshape(1000, 8) - ref).max() < 2.5e-16
    assert set(fixture["bc_type"]) == {"v  "} and len(fixture["bc_type"]) == 600


def test_vertex_classes_match_bp5_ma2(fixture):
    c = oracle.Case(10, 10, 10, nx=3, rescale=False)
    mine = c.vertex.reshape(1000, 8)
    ref = fixture["vertex"].astype(np.int64)
    assert ref.min() == 1 and ref.max() == 1331 and fixture["ma2_header"][1] == 1331
    # genmap labels vertices arbitrarily; the labelling must be a bijection of ours (same corner order)
    pairs = np.unique(np.stack([mine.ravel(), ref.ravel()], 1), axis=0)
    assert len(pairs) == 1331 and len(np.unique(pairs[:, 0])) == 1331 and len(np.unique(pairs[:, 1])) == 1331


@pytest.mark.parametrize("nx", [3, 4, 8])
def test_setvert3d_on_genmap_vertices_same_classes(fixture, nx):
    """glo_num from the reference's genmap vertex ids and from the oracle's lexicographic ids
    induce the same who-sums-with-whom structure, for several np."""
    c = oracle.Case(10, 10, 10, nx=nx, rescale=False)
    ref_vertex = np.ascontiguousarray(fixture["vertex"].astype(np.int64).ravel())
    base = classes(c.glo_num)
    for nproc in (1, 2, 32):
        g = np.zeros(c.n, dtype=np.int64)
        ngv = oracle.lib().nko_setvert3d(g, nx, 1000, ref_vertex, nproc)
        assert np.array_equal(classes(g), base)
        n = nx - 1
        assert ngv == (10 * n + 1) ** 3 - 1000 * (n - 1) ** 3 == len(np.unique(g[g > 0]))


# --------------------------------------------------------------------------- numbering
@pytest.mark.parametrize("shape,per", [((3, 2, 4), (0, 0, 0)), ((3, 3, 3), (1, 0, 0)), ((4, 3, 3), (1, 0, 1))])
def test_setvert3d_equals_coordinate_hash(shape, per):
    nx = 5
    c = oracle.Case(*shape, nx=nx, periodic=per, rescale=False, hi=tuple(float(s) for s in shape))
    # coordinate-hash numbering: nodes coincide iff coordinates coincide (mod period)
    n = nx - 1
    key = []
    for a, s, p in ((c.xm1, shape[0], per[0]), (c.ym1, shape[1], per[1]), (c.zm1, shape[2], per[2])):
        z, _ = oracle.zwgll(nx)
        # integer lattice coordinate: element index*n + local index
        k = np.rint(a * 1e6).astype(np.int64)
        if p:
            k = k % int(round(s * 1e6))
        key.append(k)
    h = (key[0] * 100_000_007 + key[1]) * 100_000_007 + key[2]
    _, inv, cnt = np.unique(h, return_inverse=True, return_counts=True)
    shared = cnt[inv] > 1
    # every shared node has a non-zero id and identical partition
    ids = c.glo_num
    assert np.all(ids[shared] > 0)
    # interior nodes are zero
    loc = np.arange(c.n) % nx ** 3
    i, j, k = loc % nx, (loc // nx) % nx, loc // nx ** 2
    interior = (i > 0) & (i < n) & (j > 0) & (j < n) & (k > 0) & (k < n)
    assert np.all(ids[interior] == 0) and np.all(ids[~interior] > 0)
    a = classes(np.where(interior, 0, inv + 1))
    assert np.array_equal(classes(ids), a)


def test_setvert3d_np_dependence_is_only_a_relabelling():
    c1 = oracle.Case(4, 4, 4, nx=4, np_ranks=1)
    c8 = oracle.Case(4, 4, 4, nx=4, np_ranks=8)
    assert not np.array_equal(c1.glo_num, c8.glo_num)  # gbtuple_rank8 buckets by mod(key,np)
    assert np.array_equal(classes(c1.glo_num), classes(c8.glo_num))
    assert c1.ngv == c8.ngv
    # vertices keep their ids (navier8.f:2052-2062)
    assert np.array_equal(c1.glo_num[:: 1][c1.glo_num <= c1.vertex.max()], c8.glo_num[c8.glo_num <= c8.vertex.max()])


# --------------------------------------------------------------------------- gs
def test_dssum_multiplicity_and_ops():
    c = oracle.Case(3, 3, 3, nx=4)
    m = c.dssum(np.ones(c.n))
    assert set(np.unique(m)) == {1.0, 2.0, 4.0, 8.0}
    assert np.array_equal(c.mult, 1.0 / m)
    rng = np.random.default_rng(2)
    u = rng.standard_normal(c.n)
    s = c.dssum(u)
    # every class carries the same value afterwards, total is conserved: sum(mult*s) == sum(u)
    assert abs((c.mult * s).sum() - u.sum()) < 1e-11
    assert np.array_equal(c.dssum(u, op=4), -c.dssum(-u, op=3))
    mx = c.dssum(u, op=4)
    assert np.all(mx >= u) and np.array_equal(c.dssum(mx, op=4), mx)
    pr = c.dssum(np.full(c.n, 2.0), op=2)
    assert np.array_equal(pr, 2.0 ** m)


# --------------------------------------------------------------------------- geometry + operator
def test_mass_and_geometry_known_answers():
    c = oracle.Case(2, 3, 2, nx=8)
    g = c.geom()
    assert abs(c.bm1().sum() - 1.0) < 1e-13  # bp5.usr:40-44 unit box
    # undeformed box: cross terms vanish, G_rr = w3 * (hy hz / hx) / ... check through A 1 = 0 and symmetry
    assert max(np.abs(g[3]).max(), np.abs(g[4]).max(), np.abs(g[5]).max()) < 1e-15
    gf = c.gf().reshape(-1, 6)
    for a, b in ((0, 0), (3, 1), (5, 2), (1, 3), (2, 4), (4, 5)):  # gf order rr,rs,rt,ss,st,tt vs core rr,ss,tt,rs,rt,st
        assert np.allclose(gf[:, a], g[b], rtol=1e-13, atol=1e-18)


@pytest.mark.parametrize("deform", [0.0, 0.05])
def test_ax_symmetric_nullspace_and_variants_agree(deform):
    c = oracle.Case(2, 2, 3, nx=8, deform=deform)
    rng = np.random.default_rng(3)
    u, v = rng.standard_normal(c.n), rng.standard_normal(c.n)
    au, pap = c.ax_bp5(u)
    av, _ = c.ax_bp5(v)
    scale = np.abs(au).max()
    assert abs((v * au).sum() - (u * av).sum()) < 1e-11 * scale * c.n ** 0.5
    assert abs(pap - (u * au).sum()) < 1e-10 * abs(pap)
    a1, _ = c.ax_bp5(np.ones(c.n))
    assert np.abs(a1).max() < 1e-12 * scale
    ah = c.axhelm(u, np.ones(c.n), np.zeros(c.n))
    assert np.abs(ah - au).max() < 1e-13 * scale
    if deform:
        assert np.abs(c.geom()[3]).max() > 1e-6
    # Helmholtz term: h2*B*u
    h2 = np.full(c.n, 0.5)
    ah2 = c.axhelm(u, np.ones(c.n), h2)
    assert np.allclose(ah2 - ah, 0.5 * c.bm1() * u, atol=1e-14 * scale)


@pytest.mark.parametrize("deform", [0.0, 0.08])
def test_setprec_is_diagonal_of_local_operator(deform):
    """setprec (hmholtz.f:380-524) is the diagonal of the local operator for constant h1, except that for
    deformed elements the reference adds the cross terms only at the 8 corners (:440-468), not along the
    12 edges -- restated as is."""
    c = oracle.Case(2, 2, 2, nx=4, deform=deform, dirichlet=(0,) * 6)
    nx = c.nx
    h1 = np.full(c.n, 1.3)
    h2 = 0.3 + 0.1 * np.sin(c.ym1)
    diag = np.zeros(c.n)
    for q in range(c.n):  # probe the local (unassembled) operator
        e = np.zeros(c.n); e[q] = 1.0
        diag[q] = c.axhelm(e, h1, h2)[q]
    g = c.geom()
    loc = np.zeros(c.n)
    oracle.lib().nko_setprec_local(loc, h1, h2, nx, c.nel, c.dt, *g[:7], None)
    l = np.arange(c.n) % nx ** 3
    ends = sum(((a == 0) | (a == nx - 1)).astype(int) for a in (l % nx, (l // nx) % nx, l // nx ** 2))
    ok = ends != 2 if deform else np.ones(c.n, bool)
    assert np.allclose(loc[ok], diag[ok], rtol=1e-12)
    if deform:
        assert not np.allclose(loc[~ok], diag[~ok], rtol=1e-6)
    assert np.allclose(c.setprec(h1, h2), 1.0 / c.dssum(loc), rtol=1e-15)


def test_manufactured_poisson_converges_spectrally():
    errs = []
    for nx in (4, 6, 8):
        c = oracle.Case(2, 2, 2, nx=nx)
        x, y, z = c.xm1, c.ym1, c.zm1
        e = np.sin(np.pi * x) * np.sin(np.pi * y) * np.sin(np.pi * z)  # bp5.usr:397-420
        f = 3 * np.pi ** 2 * e
        rhs = c.bm1() * f  # weak form; hmholtz dssum's and masks it (hmholtz.f:50-51)
        rhs = c.dssum(rhs) * c.mask
        u, it = c.cggo(rhs, np.ones(c.n), np.zeros(c.n), tin=1e-13, maxit=400, istep=100)
        errs.append(np.abs(u - e).max())
    assert errs[0] > 20 * errs[1] > 400 * errs[2] and errs[2] < 2e-6


def test_cggos_bp5_recovers_exact_solution():
    c = oracle.Case(3, 3, 3, nx=8)
    e1, r1 = c.bp5_problem()
    u, it, hist = c.cggos(r1, e1, tol=-1e-8, maxit=120, history=True)
    assert it == 120  # tol<0: fixed iteration count (bp5.par:13)
    assert oracle.glrdif(u, e1) < 1e-6
    assert np.all(hist[:, 0] > 0) and hist[-1, 3] < hist[0, 3]
    # positive tolerance exits early on max|u-x1| (bp5.usr:870)
    u2, it2 = c.cggos(r1, e1, tol=1e-3, maxit=120)
    assert 1 <= it2 < 120 and np.abs(u2 - e1).max() < 1e-3


def test_cggo_and_cggos_agree_on_solution():
    c = oracle.Case(2, 2, 2, nx=6)
    e1, r1 = c.bp5_problem()
    x, it = c.cggo(r1, np.ones(c.n), np.zeros(c.n), tin=1e-12, maxit=300, istep=100)
    assert np.abs(x - e1).max() < 1e-8 and it < 300


def test_cpu_baseline_matches_oracle_iterates():
    c = oracle.Case(2, 2, 2, nx=8)
    e1, r1 = c.bp5_problem()
    u_ref, _ = c.cggos(r1, e1, maxit=25)
    u, sec, nt = oracle.cpu_cggos(c, r1, 25, nthreads=2)
    assert sec > 0 and nt == 2
    assert np.abs(u - u_ref).max() < 1e-11 * np.abs(u_ref).max()
