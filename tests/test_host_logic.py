"""CPU tests of the host side of libnekb200: the C-ABI library loads and exports every symbol the header
declares, the host numbering (setvert3d) is bit-exact against the oracle and the reference's own BP5 mesh
fixture, and the multi-rank host logic (tuple ranking with real exchange, shared-id rendezvous) is covered with
a world_size-2/4 gloo job.  No GPU compute is invoked here."""
import os
import subprocess
import sys

import numpy as np
import pytest

import oracle

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def nb():
    import nek5000_b200 as nb
    nb.build.build_library()
    return nb


def test_library_exports_every_declared_symbol(nb):
    L = nb.lib()
    names = nb.declared_symbols()
    assert len(names) >= 60
    missing = [n for n in names if not hasattr(L, n)]
    assert not missing, missing
    for must in ("axhelm_", "cggo_", "dssum_", "dsop_", "setupds_", "fgslib_gs_setup_", "fgslib_gs_op_",
                 "fgslib_gs_op_many_", "fgslib_gs_op_fields_", "fgslib_gs_free_", "cggos_", "axhm1_", "glsc3_", "setprec_"):
        assert must in names


def test_no_gpu_is_an_error_not_a_fallback(nb):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    L = nb.lib()
    assert L.nekb_init(0, 8, 3) != 0
    assert len(L.nekb_last_error()) > 0


def test_product_never_imports_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "nek5000_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in txt and "from oracle" not in txt and "nek_oracle" not in txt, f


@pytest.mark.parametrize("dims,per,nx", [((3, 2, 2), (0, 0, 0), 5), ((2, 2, 2), (1, 0, 0), 4), ((4, 3, 2), (1, 1, 1), 3),
                                          ((1, 1, 1), (0, 0, 0), 8), ((2, 1, 1), (1, 1, 1), 6), ((5, 4, 3), (0, 0, 0), 8),
                                          ((3, 3, 3), (0, 0, 0), 2)])
@pytest.mark.parametrize("np_ranks", [1, 2, 7])
def test_setvert3d_bit_exact_vs_oracle(nb, dims, per, nx, np_ranks):
    case = oracle.Case(*dims, nx=nx, periodic=per, np_ranks=np_ranks)
    glo, ngv = nb.nek.setvert3d(nx, case.nel, case.vertex, np_ranks)
    assert np.array_equal(glo, case.glo_num)
    assert ngv == case.ngv


def test_setvert3d_on_reference_bp5_fixture(nb, golden_dir):
    """examples/bp5/bp5.ma2 vertex ids (genmap's own labelling, not lexicographic)."""
    fx = np.load(os.path.join(golden_dir, "bp5_fixture.npz"))
    vertex = fx["vertex"].astype(np.int64)
    nel = vertex.shape[0]
    for np_ranks in (1, 32):
        ref = np.zeros(512 * nel, dtype=np.int64)
        ngv_ref = oracle.lib().nko_setvert3d(ref, 8, nel, np.ascontiguousarray(vertex.reshape(-1)), np_ranks)
        glo, ngv = nb.nek.setvert3d(8, nel, vertex, np_ranks)
        assert np.array_equal(glo, ref) and ngv == ngv_ref


def _run_gloo(tmp_path, world, dims, nx):
    port = 29500 + (os.getpid() % 2000)
    out = str(tmp_path / "w")
    procs = []
    for r in range(world):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port),
                   OMP_NUM_THREADS="1")
        procs.append(subprocess.Popen([sys.executable, os.path.join(ROOT, "tests", "_gloo_worker.py"), out] +
                                      [str(d) for d in dims] + [str(nx)], env=env, cwd=ROOT,
                                      stdout=subprocess.PIPE, stderr=subprocess.STDOUT))
    logs = [p.communicate(timeout=300)[0].decode() for p in procs]
    assert all(p.returncode == 0 for p in procs), "\n".join(logs)
    return [np.load(out + f".{r}.npz") for r in range(world)]


@pytest.mark.parametrize("world,dims,nx", [(2, (3, 2, 4), 5), (4, (2, 3, 4), 4)])
def test_multirank_numbering_and_shared_ids_gloo(nb, tmp_path, world, dims, nx):
    res = _run_gloo(tmp_path, world, dims, nx)
    case = oracle.Case(*dims, nx=nx, np_ranks=world)
    nxyz = nx ** 3
    # distributed numbering == an np-rank reference run (np enters through gbtuple_rank8's mod-np buckets)
    for r in res:
        assert np.array_equal(r["glo"], case.glo_num[int(r["lo"]) * nxyz:int(r["hi"]) * nxyz])
        assert int(r["ngv"]) == case.ngv
    # shared-id lists: symmetric between the two sides of every pair and equal to the set intersection
    sets = [set(np.unique(r["glo"][r["glo"] != 0]).tolist()) for r in res]
    for a in range(world):
        peers, off, ids = res[a]["peers"], res[a]["off"], res[a]["ids"]
        expect_peers = [b for b in range(world) if b != a and sets[a] & sets[b]]
        assert peers.tolist() == expect_peers
        for k, b in enumerate(peers.tolist()):
            got = ids[off[k]:off[k + 1]]
            assert got.tolist() == sorted(sets[a] & sets[b])


def test_fast_diagonalisation_1d_systems_on_the_host(nb):
    """The 1-D generalised eigen-systems of the preconditioner setup are host code (Cholesky + Jacobi in long double): checked
    here without a GPU against the oracle restatements that tests/test_ref_pins.py pins to the reference's own gen_fast /
    hsmg_setup_fast (LAPACK dsygv): same eigenvalues, same S f(lam) S^T (eigenvector signs are free)."""
    import ctypes as C
    from oracle import hsmg
    L = nb.lib()
    P = lambda a: C.c_void_p(a.ctypes.data)
    bh, jgl, dgl = hsmg.semhat_weighted(7)
    for lbc, rbc, ll, lm, lr in ((0, 0, 0.3, 0.35, 0.4), (2, 0, 0.0, 0.5, 0.25), (0, 1, 0.2, 0.2, 0.0), (3, 2, 0.0, 1.0, 0.0), (1, 3, 0.0, 0.7, 0.0)):
        S, lam = np.zeros(64), np.zeros(8)
        assert L.nekb_fast1d_sem_host(8, lbc, rbc, ll, lm, lr, P(S), P(lam)) == 0
        So, lo = hsmg.fast1d_sem(lbc, rbc, ll, lm, lr, bh, jgl, dgl)
        assert np.abs(lam - lo).max() <= 1e-10 * np.abs(lo).max()
        S = S.reshape(8, 8)
        f = 1.0 / (1.0 + np.abs(lo))
        assert np.abs((S * f) @ S.T - (So * f) @ So.T).max() <= 1e-10 * np.abs((So * f) @ So.T).max()
    a7, b7, _, _ = hsmg.semhat(7)
    for lbc, rbc, ll, lm, lr in ((0, 0, 0.7, 1.0, 1.3), (1, 2, 0.0, 1.0, 0.0), (2, 0, 0.0, 0.4, 0.6)):
        S, lam = np.zeros(100), np.zeros(10)
        assert L.nekb_fast1d_host(7, lbc, rbc, ll, lm, lr, P(S), P(lam)) == 0
        So, lo = hsmg.fast1d(lbc, rbc, ll, lm, lr, a7, b7, 7)
        assert np.abs(lam - lo).max() <= 1e-10 * np.abs(lo).max()
        S = S.reshape(10, 10)
        f = 1.0 / (1.0 + np.abs(lo))
        assert np.abs((S * f) @ S.T - (So * f) @ So.T).max() <= 1e-10 * np.abs((So * f) @ So.T).max()
