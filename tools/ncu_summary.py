"""Summarise an .ncu-rep (ncu --set full) into the handful of numbers the design cites.
usage: python tools/ncu_summary.py gpurun_out/prof.ncu-rep [extra-regex]"""
import csv
import io
import re
import subprocess
import sys

KEYS = r"""gpu__time_duration.sum$|dram__bytes_read.sum$|dram__bytes_write.sum$|launch__registers_per_thread$|launch__grid_size|launch__block_size|
launch__occupancy_limit|sm__warps_active.avg.pct_of_peak_sustained_active|sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active|
sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active|sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active|sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active|
sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active|sm__issue_active.avg.pct_of_peak_sustained_elapsed|smsp__inst_executed.sum$|sm__inst_executed_pipe_fp64.sum$|
smsp__thread_inst_executed_per_inst_executed.ratio|smsp__sass_thread_inst_executed_op_d(fma|mul|add)_pred_on.sum$|smsp__sass_thread_inst_executed_op_f(fma|mul|add)_pred_on.sum$|
smsp__warps_eligible.avg.per_cycle_active|smsp__issue_active.avg.per_cycle_active|smsp__average_warps_issue_stalled_.*_per_issue_active|sm__cycles_elapsed.avg$|
smsp__inst_executed_op_branch|smsp__sass_average_branch_targets_threads_uniform|sm__sass_thread_inst_executed_op_.*_pred_on.sum$|smsp__pcsamp_warps_issue_stalled.*""".replace("\n", "")


def main():
    rep = sys.argv[1]
    pat = re.compile(KEYS + ("|" + sys.argv[2] if len(sys.argv) > 2 else ""))
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        print("==", r[hdr.index("Kernel Name")][:100], "grid", r[hdr.index("Grid Size")], "block", r[hdr.index("Block Size")])
        for i, h in enumerate(hdr):
            if pat.search(h) and r[i] not in ("", "0"):
                print(f"{h:95s} {units[i]:14s} {r[i]}")


if __name__ == "__main__":
    main()
