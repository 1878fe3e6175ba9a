"""Turns an .ncu-rep (ncu --set full) into the small text summary committed under profiles/.
usage: python profiles/summarize.py gpurun_out/prof.ncu-rep > profiles/<name>.txt"""
import csv
import subprocess
import sys

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smsp__inst_executed.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__block_size", "launch__grid_size",
        "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_warps", "sm__cycles_elapsed.max",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio"]


def main():
    rep = sys.argv[1]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    seen = {}
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")].split("(")[0].replace("<unnamed>::", "")
        seen.setdefault(name, []).append(r)
    print("# %s — ncu --set full --clock-control none (per launch; cold-cache, serialised replays)" % rep.split("/")[-1])
    for name, rs in seen.items():
        print("\n## %s  (%d launches captured; last one shown)" % (name, len(rs)))
        r = rs[-1]
        for w in WANT:
            if w in hdr:
                i = hdr.index(w)
                print("%-88s %16s %s" % (w, r[i], units[i]))


if __name__ == "__main__":
    main()
