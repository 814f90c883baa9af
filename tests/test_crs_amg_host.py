"""Host set-up of the aggregation hierarchy for large coarse problems (nek5000_b200/csrc/crs_amg.cuh, through the C-ABI's
nekb_crs_amg_*; no GPU involved): against the design prototype scripts/proto_coarse_amg.py (numpy / scipy) -- identical
aggregates, Galerkin identity A_c = P^T A P, and the iteration counts of CG preconditioned by a cycle over the exported levels.
The device cycle that will consume the hierarchy is not written yet."""
import os
import sys

import numpy as np
import pytest
import scipy.sparse as sp

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "scripts"))
import proto_coarse_amg as proto  # noqa: E402


@pytest.fixture(scope="module")
def nek():
    from nek5000_b200 import nek as N
    return N


def _levels(nek, A, nmax):
    C = A.tocoo()
    return nek.crs_amg_build_host(A.shape[0], C.row, C.col, C.data, nmax=nmax, theta=0.02)


def _csr(lv):
    return sp.csr_matrix((lv["val"], lv["col"], lv["rowptr"]), shape=(lv["n"], lv["n"]))


def test_hierarchy_matches_the_prototype_and_is_galerkin(nek):
    A = proto.q1_stiffness(20)                            # 8820 vertices, one Dirichlet side
    lv = _levels(nek, A, 1000)
    H = proto.Hierarchy(A, nmax=1000)
    assert [l["n"] for l in lv] == H.sizes and len(lv) == 2
    assert np.abs(_csr(lv[0]) - A).max() <= 1e-15          # duplicates summed, pattern kept
    agg, na = proto.aggregate(A, 0.02)
    assert na == lv[1]["n"] and np.array_equal(lv[0]["agg"], agg)       # integer work: identical
    P = sp.csr_matrix((np.ones(A.shape[0]), (np.arange(A.shape[0]), agg)), shape=(A.shape[0], na))
    Ac = (P.T @ A @ P).tocsr()
    assert np.abs(_csr(lv[1]) - Ac).max() <= 1e-13 * np.abs(Ac).max()
    assert np.all(np.diff(lv[1]["rowptr"]) > 0)
    for l in lv:                                           # columns ascending inside every row
        for i in (0, l["n"] // 2, l["n"] - 1):
            c = l["col"][l["rowptr"][i]:l["rowptr"][i + 1]]
            assert np.all(np.diff(c) > 0)


def test_duplicates_and_errors(nek):
    from nek5000_b200.nek import NekbError
    lv = nek.crs_amg_build_host(3, [0, 0, 1, 2, 0, 1], [0, 0, 1, 2, 1, 0], [1.0, 2.0, 4.0, 5.0, -1.0, -1.0], nmax=8)
    assert len(lv) == 1 and np.array_equal(lv[0]["rowptr"], [0, 2, 4, 5]) and np.array_equal(lv[0]["col"], [0, 1, 0, 1, 2])
    assert np.array_equal(lv[0]["val"], [3.0, -1.0, -1.0, 4.0, 5.0])
    with pytest.raises(NekbError):                          # index outside the matrix
        nek.crs_amg_build_host(3, [0, 3], [0, 0], [1.0, 1.0], nmax=8)
    # decoupled unknowns (identity rows of masked vertices) share ONE aggregate, so they cannot keep a hierarchy above nmax
    lv = nek.crs_amg_build_host(64, np.arange(64), np.arange(64), np.ones(64), nmax=8)
    assert [l["n"] for l in lv] == [64, 1] and lv[1]["val"].tolist() == [64.0] and not lv[0]["agg"].any()
    with pytest.raises(NekbError):                          # coupled, but nothing is strong at this theta: aggregation stalls
        n = 40000
        i = np.arange(n - 1)
        nek.crs_amg_build_host(n, np.concatenate([np.arange(n), i, i + 1]), np.concatenate([np.arange(n), i + 1, i]),
                               np.concatenate([np.ones(n), np.full(2 * (n - 1), -1e-3)]), nmax=8)


def test_cycle_over_the_exported_levels_converges_like_the_prototype(nek):
    """CG to 1e-13 preconditioned by one V(1,1) cycle over the library-built levels: 27 iterations at 8820 vertices (two
    levels), against 110 for Jacobi-PCG -- the numbers of scripts/proto_coarse_amg.py."""
    A = proto.q1_stiffness(20)
    lv = _levels(nek, A, 4096)
    mats = [_csr(l) for l in lv]
    Ps = [sp.csr_matrix((np.ones(l["n"]), (np.arange(l["n"]), l["agg"])), shape=(l["n"], lv[k + 1]["n"])) for k, l in enumerate(lv[:-1])]
    Ainv = np.linalg.inv(mats[-1].toarray())

    def cycle(b, l=0):
        if l == len(mats) - 1:
            return Ainv @ b
        dj = 0.7 / mats[l].diagonal()
        x = dj * b
        x = x + Ps[l] @ cycle(Ps[l].T @ (b - mats[l] @ x), l + 1)
        return x + dj * (b - mats[l] @ x)

    b = A @ np.random.default_rng(0).standard_normal(A.shape[0])
    x, it = proto.pcg(A, b, cycle)
    _, itj = proto.pcg(A, b, lambda r: r / A.diagonal())
    assert it == 27 and itj == 110
    assert np.linalg.norm(A @ x - b) <= 2e-13 * np.linalg.norm(b)


def test_smoothed_aggregation_hierarchy(nek):
    """omega_p > 0: P = (I - omega_p D^-1 A) P_tentative, A_c = P^T A P by the library's sparse products, against scipy; CG over
    the exported levels needs the prototype's 20 iterations (27 with the plain prolongation, 110 for Jacobi-PCG)."""
    A = proto.q1_stiffness(20)
    C = A.tocoo()
    lv = nek.crs_amg_build_host(A.shape[0], C.row, C.col, C.data, nmax=2048, theta=0.02, omega_p=0.66)
    assert [l["n"] for l in lv] == [8820, 343]
    n, na = lv[0]["n"], lv[1]["n"]
    T = sp.csr_matrix((np.ones(n), (np.arange(n), lv[0]["agg"])), shape=(n, na))
    Pref = (T - 0.66 * (sp.diags(1.0 / A.diagonal()) @ (A @ T))).tocsr()
    P = sp.csr_matrix((lv[0]["p_val"], lv[0]["p_col"], lv[0]["p_rowptr"]), shape=(n, na))
    assert np.abs(P - Pref).max() <= 1e-15 and 5.0 < P.nnz / n < 6.0
    assert np.abs(P @ np.ones(na) - 1.0)[np.abs(A @ np.ones(n)) < 1e-12].max() <= 1e-14      # constants are reproduced where A 1 = 0
    Ac = (Pref.T @ A @ Pref).tocsr()
    assert np.abs(_csr(lv[1]) - Ac).max() <= 1e-13 * np.abs(Ac).max()
    Ainv = np.linalg.inv(_csr(lv[1]).toarray())
    dj = 0.7 / A.diagonal()

    def cycle(b):
        x = dj * b
        x = x + P @ (Ainv @ (P.T @ (b - A @ x)))
        return x + dj * (b - A @ x)

    b = A @ np.random.default_rng(0).standard_normal(n)
    x, it = proto.pcg(A, b, cycle)
    cyc, sizes, _, _ = proto.smoothed_hierarchy_cycle(A)
    _, itp = proto.pcg(A, b, cyc)
    assert it == itp == 20 and sizes == [8820, 343]
    assert np.linalg.norm(A @ x - b) <= 2e-13 * np.linalg.norm(b)
    # the plain form still exports its (one entry per row) prolongation
    lv0 = _levels(nek, A, 2048)
    assert np.array_equal(lv0[0]["p_col"], lv0[0]["agg"]) and np.array_equal(lv0[0]["p_val"], np.ones(n))
