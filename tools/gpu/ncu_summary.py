"""Turns `ncu -i X.ncu-rep --page raw --csv` files of the observation kernels into (a) a one-page text summary per capture
(the counters DESIGN.md quotes) and (b) profiles/ncu_traffic.json, the dram traffic per launch that bench.py reports as
roofline.traffic — keyed by "<config>:<environments>".
usage: python tools/gpu/ncu_summary.py TAG config:envs=raw.csv [config:envs=raw.csv ...]"""
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
KEEP = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "smsp__inst_executed.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio", "smsp__average_warp_latency_issue_stalled_short_scoreboard.ratio",
        "smsp__average_warp_latency_issue_stalled_barrier.ratio", "smsp__average_warp_latency_issue_stalled_wait.ratio",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__occupancy_limit_shared_mem",
        "launch__occupancy_limit_registers", "dram__throughput.avg.pct_of_peak_sustained_elapsed"]
MB = {"Mbyte": 1e6, "Kbyte": 1e3, "Gbyte": 1e9, "byte": 1.0}


def main():
    tag, specs = sys.argv[1], sys.argv[2:]
    path = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    try:
        table = json.load(open(path))
    except (OSError, ValueError):
        table = {}
    lines = []
    for spec in specs:
        key, raw = spec.split("=")
        rows = list(csv.reader(open(raw)))
        hdr, units = rows[0], rows[1]
        total = 0.0
        lines.append("== %s  (%s)" % (key, os.path.basename(raw)))
        for r in rows[2:]:
            name = r[hdr.index("Kernel Name")]
            short = name.split("(")[0].replace("void <unnamed>::", "")
            lines.append("  kernel %s" % short)
            for m in KEEP:
                if m in hdr:
                    lines.append("    %-72s %s %s" % (m, r[hdr.index(m)], units[hdr.index(m)]))
            for m in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
                i = hdr.index(m)
                total += float(r[i]) * MB.get(units[i], 1.0)
        table[key] = {"dram_bytes": total, "source": "ncu --set full capture, profiles/%s_ncu_summary.txt (%s)" % (tag, key),
                      "kernels": [r[hdr.index("Kernel Name")].split("(")[0].replace("void <unnamed>::", "") for r in rows[2:]]}
        lines.append("  dram bytes per step (all observation kernels): %.1f MB" % (total / 1e6))
    with open(os.path.join(ROOT, "profiles", "%s_ncu_summary.txt" % tag), "w") as f:
        f.write("\n".join(lines) + "\n")
    json.dump(table, open(path, "w"), indent=1, sort_keys=True)
    print("\n".join(lines))


if __name__ == "__main__":
    main()
