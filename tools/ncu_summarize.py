#!/usr/bin/env python
"""Turn the ncu outputs that came back in gpurun_out/ into the tracked summaries under profiles/ (runs on the build host).

  python tools/ncu_summarize.py launches gpurun_out/r02_launches.csv profiles/r02_launches_bench.md "<command>"
  python tools/ncu_summarize.py metrics  gpurun_out/r02_prof.ncu-rep  profiles/r02_ncu_selected_metrics.csv
"""
import collections
import csv
import subprocess
import sys

METRICS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "smsp__inst_executed.sum",
           "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
           "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
           "l1tex__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
           "l1tex__m_xbar2l1tex_read_bytes.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
           "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
           "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
           "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
           "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tma.avg.pct_of_peak_sustained_active",
           "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
           "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]


def launches(src, dst, command):
    rows = list(csv.reader(l for l in open(src, errors="replace") if l.startswith('"')))
    hdr = rows[0]
    ik, im, iv = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value")
    agg = collections.OrderedDict()
    for r in rows[1:]:
        if len(r) <= iv or r[im] != "gpu__time_duration.sum":
            continue
        name = r[ik]
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += float(r[iv].replace(",", "")) / 1e3
    with open(dst, "w") as f:
        f.write("# ncu launch list\n\nCommand: `%s`\nCold-cache, serialised launches: compare SHARES with the CUDA-event numbers of the bench line, not absolutes.\n\n" % command)
        f.write("| kernel | launches | avg us | total us |\n|---|---:|---:|---:|\n")
        for name, (n, tot) in agg.items():
            f.write("| `%s` | %d | %.1f | %.0f |\n" % (name[:110], n, tot / n, tot))


def metrics(src, dst):
    raw = subprocess.check_output(["ncu", "-i", src, "--page", "raw", "--csv"]).decode(errors="replace")
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    cols = [hdr.index("Kernel Name")] + [hdr.index(m) for m in METRICS if m in hdr]
    with open(dst, "w") as f:
        w = csv.writer(f)
        w.writerow([hdr[c] for c in cols])
        w.writerow([units[c] for c in cols])
        for r in rows[2:]:
            w.writerow([r[c][:120] if c == cols[0] else r[c] for c in cols])


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2], sys.argv[3], sys.argv[4] if len(sys.argv) > 4 else "")
    else:
        metrics(sys.argv[2], sys.argv[3])
