#!/usr/bin/env python
"""BP5 Poisson (N=7, FP64) throughput: python bench.py --gpus N --steps K --warmup W [--impl reference]

A step is one cggos solve (examples/bp5/bp5.usr:367-369: `maxit` fixed CG iterations of Ax + dssum + mask + the
vector updates) of the BP5 box case; the N=1 workload is BASELINE.json configs[3]'s mesh, E = 64^3 = 262,144
elements of order N=7 (weak scaling: 64^3 elements per GPU, bricks px*py*pz).  Metric: GDOF/s with the reference's
own accounting (iterations * E_global * N^3 / seconds, bp5.usr:378-383).

Prints ONE JSON line (rank 0).  `value`: inputs resident in HBM, timed with CUDA events on the library stream, max
over ranks.  `e2e`: the same solve through the Fortran-named host-buffer entry point cggos_ (pinned host arrays,
H2D of rhs/weights and D2H of the solution inside the timed region).  `roofline`: the dominant kernel (Ax), its
algorithmic bytes per launch / its mean launch time measured live with CUDA events.  `cpu_baseline`: the REFERENCE's
own cggos loop (oracle/_ref: /root/reference's Fortran transpiled to C by oracle/f77c.py, prebuilt by build()) on every
host core, on a bounded sample (kind "reference"); when that library was not shipped, the oracle's OpenMP restatement
(kind "port").  --impl reference times that CPU leg alone.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
# stdout carries exactly one JSON line: anything a library prints there (e.g. NCCL's version banner when NCCL_DEBUG is set)
# is sent to stderr by pointing fd 1 at fd 2 for the duration of the run; emit() writes the line to the real stdout.
_REAL_STDOUT = os.dup(1)
os.dup2(2, 1)


def emit(obj) -> None:
    sys.stdout.flush()
    os.write(_REAL_STDOUT, (json.dumps(obj) + "\n").encode())

LX1 = 8
NDOF = (LX1 - 1) ** 3            # DOF per element as the reference counts them (N^3, bp5.usr:380)
NXYZ = LX1 ** 3
# algorithmic words (8 B) per grid point per CG iteration
WORDS_AX = 8.0                   # read p, 6 geometric factors, write w
WORDS_AX_CG = 12.0               # fused kernel: + read/write u (x += alpha p) + read r, write p (p = r + beta p); the p read is shared
WORDS_ITER = 19.445              # + gs/mask 1.445 + x,r update 6 + weight 1 + p update 3
WORDS_ITER_EXECUTED = 12.0 + 1.445 + 3.125   # what the fused path moves: ax_cg 12 + gs 1.445 + (r, w read; r write; 1-byte code) 3.125


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, dev: int):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200",
                                       "-i", str(dev)], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        rows = [r.strip().split(",") for r in open(self.f.name) if r.strip()]
        os.unlink(self.f.name)
        sm, reasons = [], set()
        for r in rows:
            try:
                sm.append(float(r[1]))
                out["sm_max_mhz"] = float(r[2])
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4:8]):
                    if v.strip().lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        if sm:
            sm.sort()
            out["sm_mhz"] = sm[len(sm) // 2]
        out["reasons"] = sorted(reasons)
        out["samples"] = len(sm)
        return out


def cpu_leg(steps: int, warmup: int, m_cpu: int, target_s: float):
    """The oracle's OpenMP restatement of cggos (oracle/bp5_cpu.c) on a bounded sample of the workload."""
    import oracle
    case = oracle.Case(m_cpu, m_cpu, m_cpu, nx=LX1)
    _, r1 = case.bp5_problem()
    _, sec, nt = oracle.cpu_cggos(case, r1, 2)
    its = max(2, min(500, int(target_s / max(sec / 2, 1e-9))))
    for _ in range(warmup):
        oracle.cpu_cggos(case, r1, max(1, its // 4))
    tot = 0.0
    for _ in range(steps):
        _, sec, nt = oracle.cpu_cggos(case, r1, its)
        tot += sec
    gdofs = steps * its * case.nel * NDOF / tot / 1e9
    sample = (f"E={m_cpu}^3={case.nel} elements (N=7) of the same box case, {its} CG iterations per step, {steps} steps; "
              f"gcc -O3 -march=native -fopenmp restatement of bp5.usr cggos/ax_e_bp5 + shared-memory gs")
    return gdofs, nt, sample, tot / steps * 1e3


def ref_leg(steps: int, warmup: int, target_s: float):
    """The reference's own loop (oracle/_ref) on all host cores; None when the prebuilt library is absent.  Runs in a
    fresh interpreter (python -m oracle.ref_bench): the CPU arm shares neither the CUDA context nor a single symbol with
    the product library loaded in this process."""
    try:
        r = subprocess.run([sys.executable, "-m", "oracle.ref_bench", "--steps", str(steps), "--warmup", str(warmup),
                            "--target-s", str(target_s)], cwd=ROOT, capture_output=True, text=True, timeout=900)
        if r.returncode != 0:
            raise RuntimeError(r.stderr.strip().splitlines()[-1] if r.stderr.strip() else f"exit code {r.returncode}")
        d = json.loads(r.stdout.strip().splitlines()[-1])
        return d["gdofs"], d["cores"], d["sample"], d["ms_per_step"]
    except Exception as ex:
        print(f"bench.py: reference leg unavailable ({ex}); falling back to the OpenMP port", file=sys.stderr)
        return None


def parity_check(rank, world, local, layout):
    """N > 1: before anything is timed, a small brick-partitioned BP5 problem (2x2x1 deformed elements per rank, 40 CG
    iterations) is solved on the same ranks, through the same exchange path, and every rank's part is compared with the
    oracle's UNDIVIDED solve (the checker, not the thing measured; tests/_mgpu_worker.py runs the same comparison).
    Returns {"max_rel_vs_oracle", "its", ...} (max over ranks) for the JSON line."""
    import numpy as np
    import torch
    import torch.distributed as dist
    import oracle
    from nek5000_b200 import nek
    from nek5000_b200.bp5 import BP5
    px, py, pz = layout
    nelx, nely, nelz, maxit = 2 * px, 2 * py, 1 * pz, 40
    b = BP5(nelx, nely, nelz, lx1=LX1, device=local, rank=rank, nranks=world, layout=layout, deform=0.04)
    it, _, hist = b.solve(-1e-8, maxit, history=True)
    ref = oracle.bp5_partitioned_reference(nelx, nely, nelz, layout, rank, maxit=maxit, deform=0.04)
    rel = lambda x, y: float(np.abs(x - y).max() / max(np.abs(y).max(), 1e-300))
    d = {"u": rel(b.get("u1"), ref["u"]), "r1": rel(b.get("r1"), ref["r1"]), "pap_history": rel(hist[:, 0], ref["hist"][:, 0])}
    ok = int(np.array_equal(b.get("mult"), ref["mult"]) and it == ref["it"])
    t = torch.tensor([d["u"], d["r1"], d["pap_history"], float(1 - ok)], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    u, r1, pap, bad = (float(v) for v in t.tolist())
    return {"max_rel_vs_oracle": max(u, r1, pap), "its": int(it), "its_oracle": int(ref["it"]), "solution": u, "rhs": r1,
            "pap_history": pap, "multiplicity_and_count_identical": bad == 0.0,
            "problem": f"{nelx}x{nely}x{nelz} deformed elements in bricks {px}x{py}x{pz}, {maxit} CG iterations, every rank against the undivided oracle solve"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--m", type=int, default=64, help="elements per direction per GPU (64 -> E=262,144 per GPU)")
    ap.add_argument("--maxit", type=int, default=500, help="CG iterations per solve (bp5.par:13-15)")
    ap.add_argument("--m-cpu", type=int, default=32, help="CPU-baseline sample: elements per direction")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak: --m^3 elements PER GPU (the driver's contract); strong: --m^3 elements in TOTAL, split into bricks "
                         "(BASELINE configs[3] / SURVEY 8d: E = 262,144 on 1/2/4/8 GPUs)")
    ap.add_argument("--no-parity", action="store_true", help="skip the small multi-rank parity problem run before the timed region")
    ap.add_argument("--no-general", action="store_true", help="skip the extra timing of the general-geometry operator kernel")
    ap.add_argument("--no-check", action="store_true", help="skip the fused-vs-kernel-per-statement answer check at bench scale")
    ap.add_argument("--no-e2e", action="store_true")
    a = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    warmup = max(a.warmup, 3) if a.impl == "b200" else a.warmup

    from nek5000_b200.bp5 import brick_layout
    px, py, pz = brick_layout(max(a.gpus, 1))
    if a.scaling == "strong":
        gx = gy = gz = a.m                                         # the whole job is m^3 elements
        assert a.m % px == 0 and a.m % py == 0 and a.m % pz == 0, "strong scaling: --m must be divisible by the brick counts"
        per = f"E={a.m}^3={a.m ** 3} elements in total, {a.m ** 3 // max(a.gpus, 1)} per GPU"
    else:
        gx, gy, gz = a.m * px, a.m * py, a.m * pz
        per = f"E={a.m}^3={a.m ** 3} elements per GPU"
    E_per = gx * gy * gz // max(a.gpus, 1)
    config = {"workload": f"BP5 box mesh, N=7 (lx1=8), FP64, {per} "
                          f"({gx}x{gy}x{gz} global, bricks {px}x{py}x{pz}), cggos {a.maxit} fixed CG "
                          f"iterations per step (bp5.par), identity preconditioner, all-Dirichlet box [0,1]^3",
              "elements_global": gx * gy * gz, "iterations_per_step": a.maxit,
              "l2": f"inputs larger than L2 (each n-vector {E_per * 512 * 8 / 1e9:.2f} GB, factors {E_per * 512 * 48 / 1e9:.2f} GB per GPU)"}

    if a.impl == "reference":
        if rank != 0:
            return
        leg, kind = ref_leg(max(a.steps, 1), a.warmup, target_s=5.0), "reference"
        if leg is None:
            leg, kind = cpu_leg(max(a.steps, 1), a.warmup, a.m_cpu, target_s=4.0), "port"
        gd, nt, sample, ms = leg
        # the CPU arm runs a BOUNDED SAMPLE of the workload, not the 64^3-per-GPU mesh: say so where the config is read
        config = dict(config, workload_of_the_gpu_arm=config["workload"],
                      workload="CPU arm, bounded sample of the same BP5 case (N=7, FP64, cggos, same iteration body): " + sample,
                      same_config_as_gpu_arm=False,
                      note="the value is a host-CPU throughput on this box and does not depend on --gpus")
        config.pop("elements_global", None)
        emit(({"impl": "reference", "metric": "BP5 Poisson GDOF/s (N=7, FP64)", "value": gd, "unit": "GDOF/s",
                          "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup, "ms_per_step": ms,
                          "higher_is_better": True, "scaling": a.scaling, "vs_baseline": None, "dtype": "f64",
                          "data": "synthetic", "config": config,
                          "cpu_baseline": {"value": gd, "unit": "GDOF/s", "cores": nt, "kind": kind, "sample": sample},
                          "e2e": {"value": gd, "unit": "GDOF/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))
        return

    import numpy as np
    import torch
    import torch.distributed as dist
    from nek5000_b200 import lib, nek
    from nek5000_b200._lib import check
    from nek5000_b200.bp5 import BP5
    import ctypes as C

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product has no CPU path")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    nek.init(local, LX1, 3)
    if world > 1:
        nek.comm_init_torch()
    assert world == max(a.gpus, 1), f"--gpus {a.gpus} but WORLD_SIZE={world}"
    parity = None
    if world > 1 and not a.no_parity:
        parity = parity_check(rank, world, local, (px, py, pz))
    case = BP5(gx, gy, gz, lx1=LX1, device=local, rank=rank, nranks=world, layout=(px, py, pz))
    n, E_glob = case.n, case.nel_global
    L = lib()
    try:  # how the inter-rank part of gs_op runs (0 single rank, 1 NCCL send/recv, 2 peer memory over NVLink)
        config["gs_exchange"] = {0: "none (one rank)", 1: "pack + ncclSend/ncclRecv + unpack",
                                 2: "peer-memory stores over NVLink (CUDA IPC) + epoch flags"}[int(L.nekb_gs_exchange_mode(case.gs_handle))]
    except Exception:
        pass

    def barrier():
        torch.cuda.synchronize()
        check(L.nekb_sync())
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- device-resident timing ---------------------------------------------------------------------------------
    for _ in range(warmup):
        case.solve(-1e-8, a.maxit)
    barrier()
    sampler = ClockSampler(local) if rank == 0 else None
    check(L.nekb_prof_enable(1))
    nek.launch_count(reset=True)
    dev_s, t0 = 0.0, time.perf_counter()
    for _ in range(a.steps):
        it, sec = case.solve(-1e-8, a.maxit)
        assert it == a.maxit
        dev_s += sec
    barrier()
    wall_s = time.perf_counter() - t0
    launches = nek.launch_count()
    clocks = sampler.stop() if sampler else None
    relerr = case.relerr()
    prof = {}
    for k in ("ax", "gs", "update", "pupdate"):
        s_, c_ = C.c_double(0), C.c_int64(0)
        check(L.nekb_prof_get(k.encode(), C.byref(s_), C.byref(c_)))
        prof[k] = (s_.value, c_.value)
    check(L.nekb_prof_enable(0))
    dev_s = max_over_ranks(dev_s)
    wall_s = max_over_ranks(wall_s)
    iters = a.steps * a.maxit
    value = iters * E_glob * NDOF / dev_s / 1e9
    # Which operator kernel ran: on a mesh of affine elements (this box) the library replaces the per-node factors by six
    # constants per element (decided from the registered factors, ax.cuh ax_affine_ensure).  The same solve is timed once more
    # with that switched off, so the line also carries the number for general (deformed) geometry.
    affine = bool(L.nekb_ax_affine_active()) if hasattr(L, "nekb_ax_affine_active") else False
    general = None
    if affine and not a.no_general:
        os.environ["NEKB_AX_AFFINE"] = "0"
        case.solve(-1e-8, a.maxit)
        barrier()
        check(L.nekb_prof_enable(1))
        g_s = 0.0
        for _ in range(max(1, min(a.steps, 2))):
            it, sec = case.solve(-1e-8, a.maxit)
            g_s += sec
        barrier()
        gprof = {}
        for k in ("ax", "gs", "update", "pupdate"):
            s_, c_ = C.c_double(0), C.c_int64(0)
            check(L.nekb_prof_get(k.encode(), C.byref(s_), C.byref(c_)))
            gprof[k] = s_.value / max(c_.value, 1) * 1e3
        check(L.nekb_prof_enable(0))
        del os.environ["NEKB_AX_AFFINE"]
        g_s = max_over_ranks(g_s)
        gi = max(1, min(a.steps, 2)) * a.maxit
        general = {"value": gi * E_glob * NDOF / g_s / 1e9, "unit": "GDOF/s", "kernel_ms_per_iteration": gprof,
                   "ax_roofline_frac": WORDS_AX_CG * 8 * NXYZ * case.nel / (gprof["ax"] * 1e-3) / 1e9 / peaks()[0] if gprof["ax"] > 0 else None,
                   "what": "the same solve with NEKB_AX_AFFINE=0: six factors streamed per node (ax_cg_mma_kernel, 12 words per point), "
                           "as for deformed elements"}

    # ---- the answer at bench scale: two independent code paths must agree ------------------------------------------------
    # ("bench.py asserts nothing about the answer at bench scale".)  The oracle cannot run 500 iterations on
    # 64^3 elements in the time a bench has, so the check is internal but independent: the same solve through the
    # kernel-per-statement path (NEKB_CG_FUSED=0: stand-alone Ax kernel with per-node factors, gs_op in place, separate
    # update / direction kernels -- the path the per-apply parity tests pin to the reference) against the fused path's
    # solution and relerr.
    answer = None
    if not a.no_check:
        u_fused = torch.empty(n, dtype=torch.float64, device="cuda")
        check(L.nekb_d2d(u_fused.data_ptr(), case.devptr("u1"), n * 8))
        torch.cuda.synchronize()
        os.environ["NEKB_CG_FUSED"] = "0"
        case.solve(-1e-8, a.maxit)
        del os.environ["NEKB_CG_FUSED"]
        barrier()
        u_stock = torch.empty(n, dtype=torch.float64, device="cuda")
        check(L.nekb_d2d(u_stock.data_ptr(), case.devptr("u1"), n * 8))
        torch.cuda.synchronize()
        d = float((u_fused - u_stock).abs().max() / u_stock.abs().max())
        relerr_stock = case.relerr()
        d = max_over_ranks(d)
        answer = {"max_rel_diff_fused_vs_kernel_per_statement_path": d, "relerr_fused": relerr, "relerr_kernel_per_statement": relerr_stock,
                  "iterations": a.maxit, "ok": bool(d <= 1e-9 and abs(relerr - relerr_stock) <= 1e-9)}
        assert answer["ok"], answer

    # ---- end to end through cggos_ with host buffers -----------------------------------------------------------------
    e2e = None
    if not a.no_e2e:
        pin = lambda: torch.empty(n, dtype=torch.float64, pin_memory=True)
        u1_t, rhs_t, x1_t, mult_t, binv_t = pin(), pin(), pin(), pin(), pin()
        u1, rhs, x1, mult, binv = (t.numpy() for t in (u1_t, rhs_t, x1_t, mult_t, binv_t))
        rhs[:] = case.get("r1")
        x1[:] = case.get("e1")
        mult[:] = case.get("mult")
        binv[:] = 1.0
        nek.set_ifield(1)
        nek.cggos(u1, rhs, x1, mult, binv, -1e-8, min(a.maxit, 20), "bp5")  # warm the staging buffers
        barrier()
        t0 = time.perf_counter()
        for _ in range(a.steps):
            itn = nek.cggos(u1, rhs, x1, mult, binv, -1e-8, a.maxit, "bp5")
            assert itn == a.maxit
        barrier()
        e2e_s = max_over_ranks(time.perf_counter() - t0)
        e2e = {"value": iters * E_glob * NDOF / e2e_s / 1e9, "unit": "GDOF/s", "h2d_bytes_per_step": 2 * n * 8 * world,
               "d2h_bytes_per_step": n * 8 * world, "ms_per_step": e2e_s / a.steps * 1e3,
               "api": "cggos_(u1,rhs1,x1,rmult,binv,tin,maxit,'bp5') with pinned host arrays"}

    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return

    peak, peak_src = peaks()
    ax_s, ax_n = prof["ax"]
    fused = os.environ.get("NEKB_CG_FUSED", "1") != "0"
    ax_words = WORDS_AX_CG if fused else WORDS_AX
    ax_name = ("ax_cg_mma_kernel<6,1> (u += alpha p; p = r + beta p; w = A p; pap -- 12 words/pt; one warp per element, the "
               "in-plane contractions on DMMA)" if fused
               else "ax_tma_kernel<8,3,2> (w = A p with fused pap -- 8 words/pt)")
    if affine and fused:
        ax_words = WORDS_AX_CG - 6.0
        ax_name = ("ax_cg_affine_mma_kernel<8,2> (u += alpha p; p = r + beta p; w = A p; pap; the six factors rebuilt from one "
                   "64-byte record per element: 6 words/pt + 64 B/element; one warp per element, the in-plane contractions "
                   "on mma.sync.m8n8k4.f64)")
    ax_bytes = ax_words * 8 * NXYZ * case.nel                      # per launch (this rank's elements)
    ax_gbs = ax_bytes / (ax_s / max(ax_n, 1)) / 1e9 if ax_s > 0 else None
    iter_gbs = WORDS_ITER * 8 * NXYZ * case.nel * iters / dev_s / 1e9   # whole iteration, per GPU
    tot_prof = sum(v[0] for v in prof.values())
    out = {
        "metric": "BP5 Poisson GDOF/s (N=7, FP64)", "value": value, "unit": "GDOF/s", "n_gpus": a.gpus, "steps": a.steps,
        "warmup": warmup, "ms_per_step": dev_s / a.steps * 1e3, "higher_is_better": True, "scaling": a.scaling,
        "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config,
        "relerr": relerr, "wall_ms_per_step": wall_s / a.steps * 1e3, "gpu_launches": launches, "clocks": clocks,
        "e2e": e2e, "parity": parity, "answer_check": answer, "operator_kernel": "affine elements: per-element constants" if affine else "general: per-node factors",
        "general_geometry": general,
        "affine_check": {"max_relative_deviation_of_the_registered_factors": float(L.nekb_ax_affine_deviation()),
                         "accepted_up_to": 3e-12} if hasattr(L, "nekb_ax_affine_deviation") else None,
        "roofline": {"bound": "hbm", "kernel": ax_name, "achieved": ax_gbs, "peak": peak,
                     "unit": "GB/s", "frac": (ax_gbs / peak) if ax_gbs else None, "traffic": None, "peak_source": peak_src,
                     "algorithmic_bytes_per_launch": ax_bytes, "mean_launch_ms": ax_s / max(ax_n, 1) * 1e3,
                     "share_of_step": ax_s / tot_prof if tot_prof > 0 else None,
                     "kernel_ms_per_iteration": {k: v[0] / max(v[1], 1) * 1e3 for k, v in prof.items()},
                     "whole_iteration": {"achieved": iter_gbs, "frac": iter_gbs / peak,
                                         "algorithmic_bytes_per_element_iteration": WORDS_ITER * 8 * NXYZ,
                                         "accounting": "SURVEY 8(d): 19.445 words per point (unfused kernel sequence); a value above 1 "
                                                       "is not a bandwidth above the peak: the fused path moves fewer bytes than "
                                                       "this accounting charges -- `executed` below is the hardware utilisation",
                                         "executed": {"achieved": iter_gbs * (WORDS_ITER_EXECUTED - (6.0 if affine else 0.0)) / WORDS_ITER,
                                                      "frac": iter_gbs * (WORDS_ITER_EXECUTED - (6.0 if affine else 0.0)) / WORDS_ITER / peak,
                                                      "bytes_per_element_iteration": (WORDS_ITER_EXECUTED - (6.0 if affine else 0.0)) * 8 * NXYZ,
                                                      "accounting": "bytes the fused path needs: 16.57 words per point (10.57 when the "
                                                                    "factors are rebuilt from per-element constants)"}}},
    }
    # dram__bytes_read+write per launch of the same kernel from the committed ncu --set full captures (E = 262,144)
    if case.nel == 262144 and fused:
        fn, key = (("r2t_ncu_traffic.json", "ax_cg_affine_mma_kernel<8, 2, 0>") if affine
                   else ("r2t_ncu_traffic.json", "ax_cg_mma_kernel<6, 1, 0>"))
        pj = os.path.join(ROOT, "profiles", fn)
        if os.path.exists(pj):
            try:
                out["roofline"]["traffic"] = json.load(open(pj)).get(key)
                out["roofline"]["traffic_source"] = f"profiles/{fn}"
            except Exception:
                pass
    if affine and fused:
        out["roofline"]["note"] = ("round 2 moved the two in-plane contractions of this kernel to FP64 tensor-core instructions "
                                   "(DMMA.8x8x4) after ncu showed the previous form bound by the shared-memory pipe (71 %) and "
                                   "not by DRAM (59 %): profiles/r2o_ncu_summary.md (before), profiles/r2t_ncu_summary.md (after)")
    if a.gpus == 1 and not a.no_cpu:
        leg, kind = ref_leg(2, 1, target_s=5.0), "reference"
        if leg is None:
            leg, kind = cpu_leg(2, 1, a.m_cpu, target_s=6.0), "port"
        gd, nt, sample, _ = leg
        out["cpu_baseline"] = {"value": gd, "unit": "GDOF/s", "cores": nt, "kind": kind, "sample": sample}
    emit(out)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
