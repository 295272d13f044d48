#!/usr/bin/env python
"""Summaries of an Nsight Compute report for profiles/: the raw-metric CSV of the first kernel matching a name, a short
table of the numbers DESIGN.md quotes, and (for the headline kernel) the DRAM traffic per launch that bench.py reports as
`roofline.traffic`.

usage: ncu_summary.py <report.ncu-rep> <kernel-substring> <out-prefix> [--traffic profiles/r02_traffic.json]
writes <out-prefix>_raw.csv and <out-prefix>_details.txt
"""
import csv
import json
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "smsp__inst_executed.sum", "sm__issue_active.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__warps_active.avg.per_cycle_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]


def to_bytes(value, unit):
    f = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}.get(unit.strip(), None)
    return float(value.replace(",", "")) * f if f else None


def main():
    rep, name, prefix = sys.argv[1], sys.argv[2], sys.argv[3]
    traffic_path = sys.argv[sys.argv.index("--traffic") + 1] if "--traffic" in sys.argv else None
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    row = next(r for r in rows[2:] if name in r[hdr.index("Kernel Name")])
    with open(prefix + "_raw.csv", "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["metric", "unit", "value"])
        for k, u, v in zip(hdr, units, row):
            w.writerow([k, u, v])
    d = dict(zip(hdr, row))
    u = dict(zip(hdr, units))
    lines = ["kernel: " + d["Kernel Name"], "report: " + rep, ""]
    for k in KEYS:
        if k in d:
            lines.append("%-75s %s %s" % (k, d[k], u[k]))
    stalls = sorted(((float(d[k]), k) for k in hdr if k.startswith("smsp__average_warps_issue_stalled_") and k.endswith("_per_issue_active.ratio")
                     and d[k] not in ("", "n/a")), reverse=True)
    lines.append("")
    lines.append("warp stall reasons (stalled warps per issue):")
    for v, k in stalls[:8]:
        lines.append("  %-22s %.3f" % (k[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")], v))
    open(prefix + "_details.txt", "w").write("\n".join(lines) + "\n")
    print("\n".join(lines))
    if traffic_path:
        rd, wr = to_bytes(d["dram__bytes_read.sum"], u["dram__bytes_read.sum"]), to_bytes(d["dram__bytes_write.sum"], u["dram__bytes_write.sum"])
        json.dump({"kernel": d["Kernel Name"], "report": rep, "dram_bytes_read": rd, "dram_bytes_write": wr,
                   "what": "dram__bytes_read.sum / dram__bytes_write.sum of one launch, ncu --set full of the bench workload"},
                  open(traffic_path, "w"), indent=1)


if __name__ == "__main__":
    main()
