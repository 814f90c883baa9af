"""torchrun worker: BP5 on N GPUs (one rank per GPU, NCCL) against the single-domain oracle.
Checks, per rank: numbering classes consistent with the global numbering, e1/r1, the CG history (global scalars) and
the solution after `maxit` iterations, all within 1e-10 of the oracle's global solve.  Prints 'MGPU-OK rank r'."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist
    import oracle
    from nek5000_b200 import nek
    from nek5000_b200.bp5 import BP5, brick_layout

    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    nek.init(local, 8, 3)
    nek.comm_init_torch()
    px, py, pz = brick_layout(world)
    mx, my, mz = 2, 2, 1
    nelx, nely, nelz = mx * px, my * py, mz * pz
    deform = 0.04
    b = BP5(nelx, nely, nelz, lx1=8, device=local, rank=rank, nranks=world, layout=(px, py, pz), deform=deform)
    maxit = 40
    from nek5000_b200 import lib as _lib
    mode = _lib().nekb_gs_exchange_mode(b.gs_handle)
    want = 1 if os.environ.get("NEKB_GS_P2P", "1") == "0" else 2     # peer-memory exchange over NVLink unless switched off
    assert mode == want, f"gs exchange mode {mode}, expected {want}"
    it, sec, hist = b.solve(-1e-8, maxit, history=True)
    u = b.get("u1")

    # ---- oracle: the undivided mesh; every rank draws the seed-1 ran1 stream over its local nodes (navier5.f:2665-2674)
    case = oracle.Case(nelx, nely, nelz, nx=8, deform=deform)
    nxyz = 512
    owner = np.zeros(case.nel, dtype=np.int64)
    local_of = np.zeros(case.nel, dtype=np.int64)
    lx, ly, lz = nelx // px, nely // py, nelz // pz
    for eg in range(case.nel):
        ex, ey, ez = eg % nelx, (eg // nelx) % nely, eg // (nelx * nely)
        r = (ex // lx) + px * ((ey // ly) + py * (ez // lz))
        owner[eg] = r
        local_of[eg] = (ex % lx) + lx * ((ey % ly) + ly * (ez % lz))
    rnd = np.zeros(case.n)
    nloc = lx * ly * lz * nxyz
    stream = np.zeros(nloc)
    oracle.lib().nko_rand_fld(stream, nloc)
    for eg in range(case.nel):
        rnd[eg * nxyz:(eg + 1) * nxyz] = stream[local_of[eg] * nxyz:(local_of[eg] + 1) * nxyz]
    e1 = case.dssum(rnd) * case.mult * case.mask
    ap, _ = case.ax_bp5(e1)
    r1 = case.dssum(ap) * case.mask
    uref, itref, href = case.cggos(r1, e1, maxit=maxit, history=True)

    mine = np.flatnonzero(owner == rank)
    order = mine[np.argsort(local_of[mine])]
    take = (order[:, None] * nxyz + np.arange(nxyz)[None, :]).reshape(-1)
    rel = lambda a, c: np.abs(a - c).max() / max(np.abs(c).max(), 1e-300)
    assert it == itref == maxit
    assert rel(b.get("e1"), e1[take]) <= 1e-12, "e1"
    assert rel(b.get("r1"), r1[take]) <= 1e-10, "r1"
    assert np.array_equal(b.get("mult"), case.mult[take]), "mult"
    assert rel(hist[:, 0], href[:, 0]) <= 1e-10, "pap history"
    assert np.all(np.abs(hist[:, 1] - href[:, 2]) <= 1e-8 * np.abs(href[:, 2]) + 1e-300), "(r,z) history"
    assert rel(u, uref[take]) <= 1e-10, "solution"
    # numbering: same equivalence classes as the global numbering restricted to this rank
    g = b.get("glo_num")
    gref = case.glo_num[take]
    _, inv_a = np.unique(g, return_inverse=True)
    _, inv_b = np.unique(gref, return_inverse=True)
    first_a = {}
    first_b = {}
    ca = np.array([first_a.setdefault(v, i) for i, v in enumerate(inv_a)])
    cb = np.array([first_b.setdefault(v, i) for i, v in enumerate(inv_b)])
    assert np.array_equal(ca, cb), "numbering classes"
    # ---- ophinv on N ranks: the fused 3-right-hand-side PCG (hcg.cuh) with its NCCL reductions and gs exchange against the
    # oracle's three undivided cggo solves (which tests/test_ref_pins.py pins to the reference bit for bit)
    from nek5000_b200._lib import check, lib
    L = lib()
    rng = np.random.default_rng(11)
    h1g, h2g = 1.0 + 0.3 * rng.random(case.n), 20.0 * (5.0 + rng.random(case.n))
    rhsg = [case.bm1() * rng.standard_normal(case.n) for _ in range(3)]
    vol = float(case.bm1().sum())
    nek.set_ifield(1)
    nek.set_field_handle(1, b.gs_handle)
    nek.set_step_info(20, vol)
    nek.set_param(22, 0.0)
    D = nek.DevArray
    h1d, h2d, binvd = D.from_host(h1g[take]), D.from_host(h2g[take]), D.from_host(case.binv()[take])
    maskp, multp = b.devptr("mask"), b.devptr("mult")
    itv = np.zeros(3, dtype=np.int32)
    worst = 0.0
    for tol, maxit3, ftol in ((-1e-30, 15, 1e-10), (1e-8, 300, 1e-7)):
        outs = [D(b.n) for _ in range(3)]
        rh = [D.from_host(r[take]) for r in rhsg]
        check(L.nekb_ophinv_dev(outs[0].ptr, outs[1].ptr, outs[2].ptr, rh[0].ptr, rh[1].ptr, rh[2].ptr, h1d.ptr, h2d.ptr,
                                maskp, maskp, maskp, multp, binvd.ptr, tol, maxit3, itv.ctypes.data, None))
        for k in range(3):
            f = case.dssum(rhsg[k]) * case.mask
            assert rel(rh[k].to_host(), f[take]) <= 1e-12, "ophinv rhs"
            xo, ito = case.cggo(f, h1g, h2g, tin=tol, maxit=maxit3, istep=20)
            assert itv[k] == ito, ("ophinv iteration count", k, itv[k], ito)
            d = rel(outs[k].to_host(), xo[take])
            worst = max(worst, d) if tol < 0 else worst
            assert d <= ftol, ("ophinv solution", k, d)
    # ---- a user handle in which one rank shares NOTHING while the others do (gs_p2p_setup made a different number
    # of collective calls on such a rank).  Ranks 0..world-2 share ids 1..32, the last rank holds ids of its own; two
    # handles and two exchanges back to back so that a mismatched collective would pair up wrongly and hang / corrupt.
    if world >= 3:
        for rep in range(2):
            q = np.arange(32, dtype=np.int64)
            ids = np.concatenate([1 + q if rank < world - 1 else 10_000 + 100 * rank + q, 5_000_000 + 1000 * rank + q])
            hu = nek.fgslib_gs_setup(ids)
            v = np.full(64, float(rank + 1))
            nek.fgslib_gs_op(hu, v, 1, 1, 0)
            want = np.full(64, float(rank + 1))
            if rank < world - 1:
                want[:32] = sum(range(1, world))
            assert np.array_equal(v, want), ("asymmetric gs handle", rank, v[:4], want[:4])
            nek.fgslib_gs_free(hu)
    print(f"MGPU-OK rank {rank} of {world}: its={it} rel(u)={rel(u, uref[take]):.2e} ophinv its={itv.tolist()} rel={worst:.1e} gs-exchange-mode={mode}", flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
