#!/usr/bin/env python
"""Turns the raw ncu outputs in gpurun_out/ into the tracked summaries under profiles/.

  python scripts/summarize_profiles.py launches <launches.csv> <out.md> "<command that was profiled>"
  python scripts/summarize_profiles.py full <report.ncu-rep> <out.md> "<command>"   (needs ncu on PATH, no GPU)
"""
import collections
import csv
import subprocess
import sys

FULL_KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__cycles_elapsed.max", "smsp__cycles_active.avg",
    "l1tex__t_bytes_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_bytes_pipe_lsu_mem_global_op_st.sum",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
]


def launches(path, out, cmd):
    rows = list(csv.reader(open(path, errors="ignore")))
    hi = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    h = rows[hi]
    kn, mv = h.index("Kernel Name"), h.index("Metric Value")
    agg = collections.OrderedDict()
    for r in rows[hi + 1:]:
        if len(r) <= mv:
            continue
        try:
            v = float(r[mv].replace(",", ""))
        except ValueError:
            continue
        a = agg.setdefault(r[kn].split("(")[0][:72], [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(v[1] for v in agg.values())
    with open(out, "w") as f:
        f.write(f"Command (after the same command exited 0 without ncu): `{cmd}`\n\n")
        f.write("Per-launch times are cold-cache and serialised under ncu: compare SHARES.\n\n")
        f.write("| kernel | launches | total ms | share | avg us |\n|---|---|---|---|---|\n")
        for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"| `{k}` | {v[0]} | {v[1] / 1e6:.3f} | {100 * v[1] / tot:.1f}% | {v[1] / v[0] / 1e3:.1f} |\n")
        f.write(f"\nTotal {tot / 1e6:.1f} ms over {sum(v[0] for v in agg.values())} launches.  `at::` kernels are "
                "bench.py's synthetic-input generators (outside the timed region).\n")


def full(path, out, cmd):
    txt = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    h = rows[0]
    with open(out, "w") as f:
        f.write(f"`ncu --set full --clock-control none --import-source on` of `{cmd}` (report: `{path}`)\n\n")
        names = [r[h.index("Kernel Name")].split("(")[0] for r in rows[2:]]
        f.write("| metric | unit | " + " | ".join(f"launch {i} `{n[:40]}`" for i, n in enumerate(names)) + " |\n")
        f.write("|---|---|" + "---|" * len(names) + "\n")
        for key in FULL_KEYS:
            if key in h:
                i = h.index(key)
                f.write(f"| {key} | {rows[1][i]} | " + " | ".join(r[i] for r in rows[2:]) + " |\n")


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](*sys.argv[2:5])
