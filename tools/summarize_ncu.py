#!/usr/bin/env python3
"""Summarise a gpurun visit into profiles/: per-launch times (ncu launch list), the
full-set metrics of k_tick that the roofline cites, and the bench lines."""
import csv
import io
import json
import subprocess
import sys

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "smsp__inst_executed.sum",
        "sm__inst_executed_pipe_lsu.sum", "smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct",
        "smsp__warp_issue_stalled_short_scoreboard_per_warp_active.pct", "smsp__warp_issue_stalled_barrier_per_warp_active.pct",
        "smsp__warp_issue_stalled_wait_per_warp_active.pct", "smsp__warp_issue_stalled_math_pipe_throttle_per_warp_active.pct",
        "smsp__warp_issue_stalled_lg_throttle_per_warp_active.pct", "smsp__warp_issue_stalled_mio_throttle_per_warp_active.pct"]


def main(tag, out):
    lines = ["# ncu summary %s" % tag, ""]
    for name in ("bench_%s.json" % tag, "bench_ref_%s.json" % tag):
        try:
            lines += ["## " + name, "```", open("gpurun_out/" + name).read().strip(), "```", ""]
        except OSError:
            pass
    try:
        rows = list(csv.reader(l for l in open("gpurun_out/launches_%s.csv" % tag) if l.startswith('"')))
        h = rows[0]
        ki, vi = h.index("Kernel Name"), h.index("Metric Value")
        agg = {}
        for r in rows[1:]:
            k = r[ki].split("(")[0][:60]
            agg.setdefault(k, []).append(float(r[vi].replace(",", "")))
        tot = sum(sum(v) for v in agg.values())
        lines += ["## launch list (ncu --metrics gpu__time_duration.sum, cold-cache, serialised)", "",
                  "| kernel | launches | total us | mean us | share |", "|---|---|---|---|---|"]
        for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
            lines.append("| %s | %d | %.1f | %.1f | %.1f%% |" % (k, len(v), sum(v) / 1e3, sum(v) / len(v) / 1e3, 100 * sum(v) / tot))
        lines.append("")
    except OSError:
        pass
    try:
        raw = subprocess.run(["ncu", "-i", "gpurun_out/prof_%s.ncu-rep" % tag, "--page", "raw", "--csv"],
                             capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(raw)))
        h, units = rows[0], rows[1]
        lines += ["## tick kernels, ncu --set full (per launch)", ""]
        for r in rows[2:]:
            lines.append("launch id %s: %s" % (r[0], r[h.index("Kernel Name")][:50]))
            for w in WANT:
                if w in h:
                    lines.append("  %-75s %s %s" % (w, r[h.index(w)], units[h.index(w)]))
            lines.append("")
    except Exception as e:  # no report
        lines.append("(no full capture: %s)" % e)
    open(out, "w").write("\n".join(lines) + "\n")
    print("wrote", out)


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
