"""BASELINE config 3 (SURVEY.md 8d "E sweep"): BP5 CG iteration rate against the element count on one GPU.
    python scripts/bench_sweep.py [--dims 8,16,32,64,128] [--its 100]
A dims entry is m (an m^3 box) or axbxc.  Prints one JSON line: per size E, ms per iteration, GDOF/s and the fraction of
the HBM roofline by SURVEY 8(d)'s 79,648 algorithmic bytes per element-iteration (peak from MEASURED_PEAKS.json)."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--dims", default="8,16,32,64,128")
    ap.add_argument("--its", type=int, default=100)
    a = ap.parse_args()
    from nek5000_b200 import nek
    from nek5000_b200.bp5 import BP5
    try:
        peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        peak = None
    rows = []
    for d in a.dims.split(","):
        dims = tuple(int(v) for v in d.split("x")) if "x" in d else (int(d),) * 3
        nek.finalize()
        nek.init(0, 8, 3)
        try:
            import torch
            free0 = torch.cuda.mem_get_info(0)
            b = BP5(*dims, lx1=8)
            b.solve(-1e-8, 5)
            it, sec = b.solve(-1e-8, a.its)
            free1 = torch.cuda.mem_get_info(0)
        except Exception as ex:            # a size that does not fit fails loudly inside the library; the sweep records it
            rows.append({"dims": dims, "E": dims[0] * dims[1] * dims[2], "error": str(ex)[-200:]})
            print(json.dumps(rows[-1]), file=sys.stderr, flush=True)
            continue
        gbs = 79648.0 * b.nel * it / sec / 1e9
        # bytes the fused path moves: 16.57 words per point, 10.57 when the operator kernel rebuilds the factors from per-element
        # constants (decided by the library from the registered factors: nekb_ax_affine_active)
        from nek5000_b200 import lib
        affine = bool(lib().nekb_ax_affine_active())
        ex = (10.57 if affine else 16.57) * 8 * 512 * b.nel * it / sec / 1e9
        rows.append({"dims": dims, "E": b.nel, "ms_per_iteration": sec / it * 1e3, "gdofs": it * b.nel * 343 / sec / 1e9,
                     "operator_kernel": "affine (per-element constants)" if affine else "general (per-node factors)",
                     "executed_GBs": ex, "frac_of_hbm_peak_executed": ex / peak if peak else None,
                     "alg_GBs": gbs, "frac_of_hbm_peak_survey_accounting": gbs / peak if peak else None, "relerr": b.relerr(),
                     "hbm_used_GB": (free1[1] - free1[0]) / 1e9, "hbm_total_GB": free1[1] / 1e9})
        print(json.dumps(rows[-1]), file=sys.stderr, flush=True)
        del b
    nek.finalize()
    print(json.dumps({"workload": "BP5 cggos, N=7, FP64, one GPU", "its": a.its, "hbm_peak_GBs": peak, "sweep": rows}))


if __name__ == "__main__":
    main()
