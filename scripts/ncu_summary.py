"""Summarises ncu captures brought back in gpurun_out/ into profiles/ (tracked).

    python scripts/ncu_summary.py <tag> <launches.csv> <rep1.ncu-rep> [<rep2.ncu-rep> ...]
"""
import collections
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__inst_executed.sum", "lts__t_sector_hit_rate.pct",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sectors_srcunit_tex_op_read.sum", "lts__t_sectors_srcunit_tex_op_write.sum",
        "sm__cycles_elapsed.max"]


def short(name):
    name = name.replace("nekb::", "").replace("void ", "")
    return name.split("(")[0]


def main():
    tag, launches = sys.argv[1], sys.argv[2]
    reps = sys.argv[3:]
    out = [f"# ncu summary {tag}", ""]
    if os.path.exists(launches):
        rows = [r for r in csv.reader(open(launches)) if len(r) > 10]
        hdr = rows[0]
        if "Kernel Name" in hdr:
            ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
            data = rows[1:]
        else:                       # a tail of the file (the whole list is too large to bring back): ncu's fixed column order
            ki, vi = 4, 14
            data = rows
        # the last cggos solve = the timed step: everything after the last init kernel
        marker = os.environ.get("NCU_STEP_MARKER", "cggos_init")   # first kernel of a solve (ophinv: hcg_prep_kernel)
        last = max((i for i, r in enumerate(data) if marker in r[ki]), default=0)
        agg, cnt = collections.OrderedDict(), collections.Counter()
        for r in data[last:]:
            k = short(r[ki])
            agg[k] = agg.get(k, 0.0) + float(r[vi].replace(",", ""))
            cnt[k] += 1
        tot = sum(agg.values())
        out += [f"## Launch list of the timed step (`ncu --metrics gpu__time_duration.sum --clock-control none`, file {os.path.basename(launches)})",
                "", "Per-launch times are cold-cache and serialised; compare SHARES.", "",
                "| kernel | launches | total ms | mean ms | share |", "|---|---:|---:|---:|---:|"]
        for k, v in agg.items():
            out.append(f"| `{k}` | {cnt[k]} | {v / 1e6:.3f} | {v / 1e6 / cnt[k]:.4f} | {100 * v / tot:.1f} % |")
        out += ["", f"All launches of the run: {len(data)}; setup kernels (before the first solve): "
                    f"{', '.join(sorted({short(r[ki]) for r in data[:last]} - set(agg)))}", ""]
    traffic = {}
    for rep in reps:
        if rep.endswith(".csv"):      # `ncu -i x.ncu-rep --page raw --csv > x.raw.csv` done on the GPU box (reports are 15 MB each)
            txt = open(rep).read()
        else:
            txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(txt.splitlines()))
        hdr, units = rows[0], rows[1]
        out += [f"## `ncu --set full --clock-control none` : {os.path.basename(rep)}", ""]
        for r in rows[2:]:
            name = short(r[hdr.index("Kernel Name")])
            out += [f"### `{name}`", "", "| metric | value | unit |", "|---|---:|---|"]
            vals = {}
            for w in WANT:
                if w in hdr:
                    i = hdr.index(w)
                    out.append(f"| {w} | {r[i]} | {units[i]} |")
                    vals[w] = (float(r[i].replace(",", "")), units[i])
            out.append("")
            if "dram__bytes_read.sum" in vals:
                scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}
                rb = vals["dram__bytes_read.sum"][0] * scale[vals["dram__bytes_read.sum"][1]]
                wb = vals["dram__bytes_write.sum"][0] * scale[vals["dram__bytes_write.sum"][1]]
                traffic.setdefault(name, []).append(rb + wb)
    os.makedirs(os.path.join(ROOT, "profiles"), exist_ok=True)
    open(os.path.join(ROOT, "profiles", f"{tag}_ncu_summary.md"), "w").write("\n".join(out) + "\n")
    tj = {k: sum(v) / len(v) for k, v in traffic.items()}
    json.dump(tj, open(os.path.join(ROOT, "profiles", f"{tag}_ncu_traffic.json"), "w"), indent=1)
    print("\n".join(out[:40]))


if __name__ == "__main__":
    main()
