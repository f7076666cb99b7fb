"""Key metrics of an ncu report, one row per kernel.  usage: python profiles/ncu_summary.py <file.ncu-rep>"""
import csv
import subprocess
import sys

WANT = [
    ("gpu__time_duration.sum", "dur"),
    ("dram__bytes_read.sum", "dram_rd"),
    ("dram__bytes_write.sum", "dram_wr"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm%"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue%"),
    ("smsp__inst_executed.sum", "warp_inst"),
    ("smsp__thread_inst_executed_per_inst_executed.ratio", "thr/inst"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ%"),
    ("launch__registers_per_thread", "regs"),
    ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "st_long"),
    ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "st_short"),
    ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "st_bar"),
    ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "st_wait"),
    ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "st_math"),
    ("smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "st_mio"),
    ("smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", "st_lg"),
    ("smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio", "st_br"),
    ("smsp__average_warps_issue_stalled_membar_per_issue_active.ratio", "st_membar"),
    ("smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio", "st_noinst"),
]


def main(path):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    h, units = rows[0], rows[1]
    ik = h.index("Kernel Name")
    print("kernel," + ",".join(f"{s}[{units[h.index(m)]}]" if m in h else s for m, s in WANT))
    for r in rows[2:]:
        vals = []
        for m, _ in WANT:
            vals.append(r[h.index(m)] if m in h else "")
        print(f"\"{r[ik]}\"," + ",".join(vals))


if __name__ == "__main__":
    main(sys.argv[1])
