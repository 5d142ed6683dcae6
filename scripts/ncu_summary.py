"""Summarise an .ncu-rep (raw page) into a short per-kernel table. Usage: python scripts/ncu_summary.py file.ncu-rep"""
import csv
import subprocess
import sys

WANT = [
    ("gpu__time_duration.sum", "time"),
    ("dram__bytes_read.sum", "dram_rd"),
    ("dram__bytes_write.sum", "dram_wr"),
    ("lts__t_bytes.sum", "l2_bytes"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_pct"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm_pct"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue_pct"),
    ("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "fma_pipe_pct"),
    ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "lsu_pct"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "smem_wavefront_pct"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "occupancy_pct"),
    ("launch__registers_per_thread", "regs"),
    ("launch__occupancy_limit_registers", "occ_lim_regs"),
    ("launch__occupancy_limit_shared_mem", "occ_lim_smem"),
    ("smsp__inst_executed.sum", "warp_inst"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem_bank_conflicts"),
    ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall_long_sb"),
    ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stall_short_sb"),
    ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "stall_barrier"),
    ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "stall_math_throttle"),
    ("smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "stall_mio"),
    ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "stall_wait"),
    ("smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "stall_not_selected"),
    ("smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio", "stall_dispatch"),
    ("smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", "stall_lg"),
    ("smsp__average_warps_issue_stalled_membar_per_issue_active.ratio", "stall_membar"),
    ("smsp__average_warps_issue_stalled_sleeping_per_issue_active.ratio", "stall_sleep"),
    ("smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio", "stall_no_inst"),
    ("smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio", "stall_branch"),
]


def main(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    for r in rows[2:]:
        print("==", r[idx["Kernel Name"]][:70], " grid", r[idx.get("Grid Size", 0)], "block", r[idx.get("Block Size", 0)])
        for k, short in WANT:
            if k in idx:
                print(f"   {short:22s} {r[idx[k]]:>16s} {units[idx[k]]}")


if __name__ == "__main__":
    main(sys.argv[1])
