"""Condense `ncu -i X.ncu-rep --page raw --csv` output into the per-kernel table kept under profiles/.
Usage: python tests/gpu_checks/ncu_summary.py raw.csv out.csv ["comment line"]"""
import csv
import sys

COLS = [("time[ms]", "gpu__time_duration.sum"), ("dram_read[Gbyte]", "dram__bytes_read.sum"),
        ("dram_write[Gbyte]", "dram__bytes_write.sum"),
        ("dram_pct_of_peak[%]", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
        ("sm_pct[%]", "sm__throughput.avg.pct_of_peak_sustained_elapsed"),
        ("warps_active_pct[%]", "sm__warps_active.avg.pct_of_peak_sustained_active"),
        ("issue_active_per_cycle", "smsp__issue_active.avg.per_cycle_active"),
        ("warps_eligible_per_cycle[warp]", "smsp__warps_eligible.avg.per_cycle_active"),
        ("regs[register/thread]", "launch__registers_per_thread"),
        ("smem_dyn[Kbyte/block]", "launch__shared_mem_per_block_dynamic"),
        ("occ_lim_regs[block]", "launch__occupancy_limit_registers"),
        ("occ_lim_smem[block]", "launch__occupancy_limit_shared_mem"),
        ("tensor_pipe_active_pct[%]", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"),
        ("fma_pipe_active_pct[%]", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active"),
        ("lsu_wavefronts_pct[%]", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed"),
        ("smem_wavefronts", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"),
        ("smem_bank_conflicts", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"),
        ("inst_executed", "smsp__inst_executed.sum"),
        ("stall_long_scoreboard[inst]", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio"),
        ("stall_short_scoreboard[inst]", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio"),
        ("stall_barrier[inst]", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio"),
        ("stall_math_pipe[inst]", "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio"),
        ("stall_mio[inst]", "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio"),
        ("stall_wait[inst]", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio"),
        ("stall_not_selected[inst]", "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio")]


def main(raw, out, comment=""):
    rows = [r for r in csv.reader(open(raw)) if r]
    while rows and "Kernel Name" not in rows[0]:
        rows.pop(0)
    hdr, body = rows[0], rows[2:]
    ix = {h: i for i, h in enumerate(hdr)}
    with open(out, "w", newline="") as f:
        w = csv.writer(f)
        if comment:
            w.writerow(["# " + comment])
        w.writerow(["kernel", "grid", "block"] + [c for c, _ in COLS])
        for r in body:
            name = r[ix["Kernel Name"]].split("(")[0]
            w.writerow([name, r[ix["Grid Size"]], r[ix["Block Size"]]] + [r[ix[m]] if m in ix else "" for _, m in COLS])


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else "")
