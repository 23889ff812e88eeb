"""Summarise an .ncu-rep (raw page) into the handful of metrics DESIGN.md / profiles/ quote."""
import csv, subprocess, sys
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "launch__occupancy_limit_warps", "launch__grid_size", "launch__block_size", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_cbu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_adu.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tensor.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum", "lts__t_sectors_op_read.sum", "lts__t_sectors_op_write.sum",
        "smsp__cycles_active.avg", "sm__cycles_elapsed.max", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "smsp__pcsamp_warps_issue_stalled_long_scoreboard", "smsp__pcsamp_warps_issue_stalled_short_scoreboard",
        "smsp__pcsamp_warps_issue_stalled_wait", "smsp__pcsamp_warps_issue_stalled_math_pipe_throttle",
        "smsp__pcsamp_warps_issue_stalled_not_selected", "smsp__pcsamp_warps_issue_stalled_selected",
        "smsp__pcsamp_warps_issue_stalled_branch_resolving", "smsp__pcsamp_warps_issue_stalled_barrier",
        "smsp__pcsamp_warps_issue_stalled_no_instructions", "smsp__pcsamp_warps_issue_stalled_dispatch_stall",
        "smsp__pcsamp_warps_issue_stalled_mio_throttle", "smsp__pcsamp_warps_issue_stalled_lg_throttle",
        "smsp__pcsamp_warps_issue_stalled_imc_miss", "smsp__pcsamp_sample_count", "sm__inst_executed_pipe_tc",
        "sm__inst_executed_pipe_tmem", "sm__inst_executed_pipe_tma"]
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units, data = rows[0], rows[1], rows[2:]
name_col = hdr.index("Kernel Name")
print("kernels:", [r[name_col][:60] for r in data])
for i, h in enumerate(hdr):
    if h in KEYS or any(h == k for k in KEYS):
        print(f"{h} [{units[i]}]:", ", ".join(r[i] for r in data))
