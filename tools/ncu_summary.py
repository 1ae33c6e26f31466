"""tools/ncu_summary.py -- turn ncu output brought back in gpurun_out/ into the tracked summaries under profiles/.

    python tools/ncu_summary.py launches gpurun_out/launches_X.csv profiles/rNN_launches_X.md   [step-marker kernel]
    python tools/ncu_summary.py full     gpurun_out/prof_X.ncu-rep  profiles/rNN_full_X.md

`launches`: the --metrics gpu__time_duration.sum launch list -> per-kernel totals and shares (of the last full step
when a marker kernel name is given), plus the raw per-launch table of that step.
`full`: selected raw-page metrics of every captured launch of a --set full report.
"""
import collections
import csv
import subprocess
import sys

FULL_METRICS = [
    ("gpu__time_duration.sum", "time"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("launch__registers_per_thread", "regs"),
    ("launch__occupancy_limit_shared_mem", "occ_lim_smem"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps_active%"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm_thr%"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_thr%"),
    ("dram__bytes_read.sum", "dram_rd"),
    ("dram__bytes_write.sum", "dram_wr"),
    ("lts__t_sector_hit_rate.pct", "l2_hit%"),
    ("sm__pipe_tensor_subpipe_imma_cycles_active.avg.pct_of_peak_sustained_active", "tensor_imma%"),
    ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "alu%"),
    ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "xu%"),
    ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "lsu%"),
    ("smsp__issue_active.avg.pct", "issue%"),
    ("smsp__inst_executed.sum", "warp_insts"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem_conflicts"),
]


def read_ncu_csv(path_or_lines):
    lines = [ln for ln in path_or_lines if not ln.startswith("==")]
    return list(csv.DictReader(lines))


def launches(src, dst, marker=None):
    rows = read_ncu_csv(open(src))
    names = [r["Kernel Name"].split("(")[0].replace("void ", "") for r in rows]
    dur = [float(r["Metric Value"].replace(",", "")) / 1e3 for r in rows]  # ns -> us
    lo, hi = 0, len(rows)
    if marker:
        # a step ends with the marker kernel (e.g. k_nms_center); take the first complete step after the warm-up steps
        ends = [i for i, n in enumerate(names) if n.startswith(marker)]
        which = int(sys.argv[5]) if len(sys.argv) > 5 else 1
        if len(ends) > which:
            lo, hi = ends[which - 1] + 1, ends[which] + 1
    agg = collections.OrderedDict()
    for n, d in zip(names[lo:hi], dur[lo:hi]):
        a = agg.setdefault(n, [0, 0.0])
        a[0] += 1
        a[1] += d
    tot = sum(a[1] for a in agg.values())
    with open(dst, "w") as f:
        f.write("# ncu launch list summary (%s)\n\n" % src)
        f.write("Source: `ncu --metrics gpu__time_duration.sum --clock-control none` (cold-cache, serialised launches: compare shares).\n")
        f.write("Launches %d..%d of %d (one full step%s); total %.1f us.\n\n" % (lo, hi, len(rows), " ending with `%s`" % marker if marker else "", tot))
        f.write("| kernel | launches | total us | share | avg us |\n|---|---:|---:|---:|---:|\n")
        for n, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write("| `%s` | %d | %.1f | %.1f%% | %.1f |\n" % (n, a[0], a[1], 100 * a[1] / tot, a[1] / a[0]))
        f.write("\n## per launch\n\n| # | kernel | grid | block | us |\n|---:|---|---|---|---:|\n")
        for i in range(lo, hi):
            f.write("| %d | `%s` | %s | %s | %.1f |\n" % (i - lo, names[i], rows[i]["Grid Size"], rows[i]["Block Size"], dur[i]))
    print("wrote", dst)


def full(src, dst):
    out = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader([ln for ln in out.splitlines() if not ln.startswith("==")]))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    with open(dst, "w") as f:
        f.write("# ncu --set full summary (%s)\n\n" % src)
        f.write("One column per captured launch; values as ncu reports them (`--clock-control none`).\n\n")
        f.write("| metric | unit | " + " | ".join("#%d" % i for i in range(len(data))) + " |\n")
        f.write("|---|---|" + "---:|" * len(data) + "\n")
        f.write("| kernel | | " + " | ".join(d[idx["Kernel Name"]].split("(")[0][:24] for d in data) + " |\n")
        for m, short in FULL_METRICS:
            if m not in idx:
                continue
            f.write("| %s (`%s`) | %s | " % (short, m, units[idx[m]]) + " | ".join(d[idx[m]] for d in data) + " |\n")
    print("wrote", dst)


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2], sys.argv[3], sys.argv[4] if len(sys.argv) > 4 else None)
    else:
        full(sys.argv[2], sys.argv[3])
