#!/usr/bin/env python
"""Summarise .ncu-rep captures into small text files under profiles/.

    python tools/ncu_summary.py gpurun_out/prof_k2s.ncu-rep profiles/r01_k2_gather.txt

Reads the report with `ncu -i ... --page raw --csv` and keeps, per kernel launch, the metrics
the roofline discussion in DESIGN.md uses (duration, DRAM bytes, throughput percentages, fp64
pipe utilisation, occupancy, registers, grid).
"""
import csv
import io
import subprocess
import sys

KEEP = [
    ("gpu__time_duration.sum", "duration"),
    ("dram__bytes_read.sum", "dram_read"),
    ("dram__bytes_write.sum", "dram_write"),
    ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram_pct_of_peak"),
    ("lts__t_bytes.sum", "l2_bytes"),
    ("l1tex__t_bytes.sum", "l1_bytes"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm_pct_of_peak"),
    ("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "fp64_pipe_pct"),
    ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "lsu_pipe_pct"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved_occupancy_pct"),
    ("sm__maximum_warps_per_active_cycle_pct", "theoretical_occupancy_pct"),
    ("launch__registers_per_thread", "regs"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("launch__shared_mem_per_block_dynamic", "dyn_smem"),
    ("launch__shared_mem_per_block_static", "static_smem"),
    ("smsp__inst_executed.sum", "warp_insts"),
    ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall_long_scoreboard"),
    ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "stall_math_throttle"),
    ("smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", "stall_lg_throttle"),
    ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stall_short_scoreboard"),
    ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "stall_wait"),
    ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "stall_barrier"),
    ("smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "stall_mio_throttle"),
    ("smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio", "stall_no_instruction"),
    ("smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "stall_not_selected"),
    ("smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio", "stall_dispatch"),
    ("smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio", "stall_branch_resolving"),
    ("smsp__average_warps_issue_stalled_tex_throttle_per_issue_active.ratio", "stall_tex_throttle"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue_active_pct"),
    ("sm__inst_executed.avg.per_cycle_active", "ipc_active"),
    ("l1tex__t_sector_hit_rate.pct", "l1_hit_pct"),
    ("lts__t_sector_hit_rate.pct", "l2_hit_pct"),
    ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "alu_pipe_pct"),
    ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "fma_pipe_pct"),
    ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "xu_pipe_pct"),
    ("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "fp64_cycles_active_pct"),
]


def main():
    rep, out = sys.argv[1], sys.argv[2]
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    lines = [f"# source: {rep}  (ncu --set full --clock-control none; per-launch values)"]
    for r in data:
        lines.append("")
        lines.append("kernel: " + r[hdr.index("Kernel Name")])
        for metric, label in KEEP:
            if metric in hdr:
                i = hdr.index(metric)
                lines.append(f"  {label:28s} {r[i]} {units[i]}   [{metric}]")
    with open(out, "w") as fh:
        fh.write("\n".join(lines) + "\n")
    print("\n".join(lines))


if __name__ == "__main__":
    main()
