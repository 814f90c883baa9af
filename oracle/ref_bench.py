"""CPU timing of the REFERENCE's own BP5 loop (examples/bp5/bp5.usr:797-899 cggos with ax_e_bp5, dssum, the vector
updates), i.e. of oracle/_ref -- /root/reference's Fortran statements transpiled by oracle/f77c.py and compiled with
gcc -O2.  TEST / BENCH INFRASTRUCTURE ONLY (bench.py's `--impl reference` arm and `cpu_baseline` leg).

The reference parallelises over MPI ranks through gslib, which is not vendored and cannot be fetched; the stand-in
gather-scatter is single-rank.  To load every host core the way `mpiexec -np P` would, P independent single-rank
processes each solve their own m^3-element box at the same time (barrier, then one cggos call each, step time = the
slowest): the aggregate is what P ranks achieve WITHOUT paying for the inter-rank exchange, i.e. an upper bound for the
reference's MPI path on P cores.
"""
from __future__ import annotations

import multiprocessing as mp
import os
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
LX1, NDOF = 8, 343
LGMRES = 2          # GMRES storage is not used by BP5; a small lgmres keeps the static COMMON footprint down


def lib_path(m: int) -> str:
    return os.path.join(HERE, "_ref", f"libnekref_lx{LX1}e{m ** 3}g{LGMRES}.so")


def available(m: int = 16) -> bool:
    """Builds / refreshes the library ONCE, in the calling process (the forked workers must find it up to date: eight of
    them rebuilding the same file at once is a race).  On the GPU box, where /root/reference is absent, the prebuilt file
    is used as it is."""
    try:
        from . import ref_build
        return os.path.exists(ref_build.build(LX1, LX1, m ** 3, LGMRES))
    except Exception:
        return os.path.exists(lib_path(m)) and not os.path.isdir(os.path.join(os.environ.get("NEK_REFERENCE", "/root/reference"), "core"))


def _worker(rank, m, plan, bar, q):
    try:
        import numpy as np
        sys.path.insert(0, os.path.dirname(HERE))
        import oracle
        from oracle.ref import RefCase
        case = oracle.Case(m, m, m, nx=LX1)
        rc = RefCase(case, lelt=m ** 3, lgmres=LGMRES, fresh=False)
        R, n = rc.R, case.n
        R.var("uparam")[0:3] = (-1e-8, 1, 1)
        R.call("bp5")                                   # geodatstd, e1, r1 (+ one iteration)
        v = lambda nm: R.var(nm, "bp5").ravel(order="F")
        u1, r1, e1 = v("u1"), v("r1"), v("e1")
        vmult, binv = R.var("vmult").ravel(order="F"), R.var("binvm1").ravel(order="F")
        times = []
        for its in plan:
            bar.wait()
            t0 = time.perf_counter()
            R.call("cggos", u1, r1, e1, vmult, binv, -1e-8, int(its), "bp5")
            times.append(time.perf_counter() - t0)
        err = float(np.abs(u1[:n] - e1[:n]).max() / max(np.abs(e1[:n]).max(), 1e-300))
        q.put((rank, times, err))
    except Exception as ex:  # a dead worker must not leave the others at the barrier
        try:
            bar.abort()
        except Exception:
            pass
        q.put((rank, None, repr(ex)))


def host_cores() -> int:
    try:
        n = len(os.sched_getaffinity(0))
    except Exception:
        n = os.cpu_count() or 1
    try:  # a cgroup CPU quota below the visible core count would make P processes time-share
        quota, period = open("/sys/fs/cgroup/cpu.max").read().split()[:2]
        if quota != "max":
            n = max(1, min(n, int(float(quota) / float(period))))
    except Exception:
        pass
    return n


def run(steps: int, warmup: int, m: int = 16, nproc: int | None = None, target_s: float = 5.0, max_its: int = 500):
    """Returns dict(gdofs, cores, its, ms_per_step, sample).  Each step: every process runs `its` cggos iterations on
    its own m^3 box; its is calibrated from a 2-iteration probe so that a step takes about target_s."""
    if nproc is None:
        nproc = host_cores()
        try:
            import psutil
            nproc = max(1, min(nproc, int(psutil.virtual_memory().available // (3 << 30))))   # ~1.2 GB touched per process
        except Exception:
            pass
    if not available(m):
        raise RuntimeError("oracle/_ref is neither prebuilt nor buildable here")
    ctx = mp.get_context("fork")
    # phase 1: probe
    def launch(plan):
        bar, q = ctx.Barrier(nproc), ctx.Queue()
        ps = [ctx.Process(target=_worker, args=(r, m, plan, bar, q)) for r in range(nproc)]
        for p in ps:
            p.start()
        try:
            res = [q.get(timeout=600) for _ in ps]          # a worker that died without reporting must not hang the bench
        except Exception:
            for p in ps:
                p.kill()
            raise RuntimeError("reference worker did not report within 600 s")
        for p in ps:
            p.join(30)
        bad = [r for r in res if r[1] is None]
        if bad:
            raise RuntimeError(f"reference worker failed: {bad[0][2]}")
        return res
    probe = launch([2])
    t_it = max(r[1][0] for r in probe) / 2
    its = int(max(2, min(max_its, target_s / max(t_it, 1e-9))))
    res = launch([max(1, its // 4)] * warmup + [its] * steps)
    step_t = [max(r[1][warmup + k] for r in res) for k in range(steps)]
    tot = sum(step_t)
    E = m ** 3
    gd = steps * its * nproc * E * NDOF / tot / 1e9
    sample = (f"{nproc} independent single-rank processes (one per host core; gslib is not vendored, so no inter-rank "
              f"exchange: an upper bound for the reference's MPI path on {nproc} cores), each E={m}^3={E} elements (N=7) of "
              f"the same all-Dirichlet box, {its} cggos iterations per step, {steps} steps; reference Fortran "
              f"(bp5.usr cggos/ax_e_bp5, dssum) transpiled to C by oracle/f77c.py, gcc -O2 -ffp-contract=off")
    return dict(gdofs=gd, cores=nproc, its=its, ms_per_step=tot / steps * 1e3, sample=sample,
                relerr=max(r[2] for r in res))


if __name__ == "__main__":
    import argparse
    import json
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--warmup", type=int, default=1)
    ap.add_argument("--target-s", type=float, default=3.0)
    ap.add_argument("--nproc", type=int, default=None)
    a = ap.parse_args()
    if not available(16):
        raise SystemExit("oracle/_ref/libnekref_lx8e4096g2.so is neither prebuilt nor buildable here")
    print(json.dumps(run(steps=a.steps, warmup=a.warmup, nproc=a.nproc, target_s=a.target_s)))
