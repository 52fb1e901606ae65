"""Summarise an .ncu-rep (read here, no GPU needed): python scripts/ncu_summary.py <rep> [kernel-index]"""
import csv, subprocess, sys
KEYS = """gpu__time_duration.sum dram__bytes_read.sum dram__bytes_write.sum gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed
sm__throughput.avg.pct_of_peak_sustained_elapsed l1tex__throughput.avg.pct_of_peak_sustained_elapsed lts__throughput.avg.pct_of_peak_sustained_elapsed
sm__warps_active.avg.pct_of_peak_sustained_active launch__registers_per_thread launch__block_size launch__grid_size launch__shared_mem_per_block_dynamic launch__shared_mem_per_block_static
smsp__inst_executed.sum smsp__issue_active.avg.pct_of_peak_sustained_active sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active
sm__inst_executed_pipe_lsu.sum sm__inst_executed_pipe_fma.sum sm__inst_executed_pipe_alu.sum
l1tex__t_sector_hit_rate.pct lts__t_sector_hit_rate.pct
l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum l1tex__t_output_wavefronts_pipe_lsu_mem_global_op_ld.sum
l1tex__t_output_wavefronts_pipe_lsu_mem_global_op_ld.sum.pct_of_peak_sustained_elapsed l1tex__t_requests_pipe_lsu_mem_global_op_st.sum
l1tex__data_pipe_lsu_wavefronts.sum.pct_of_peak_sustained_elapsed l1tex__data_pipe_lsu_wavefronts_mem_shared.sum l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum
l1tex__lsu_writeback_active.avg.pct_of_peak_sustained_elapsed sm__cycles_elapsed.avg
smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio
smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio
smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio
smsp__average_warps_issue_stalled_wait_per_issue_active.ratio smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio""".split()
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
h, u = rows[0], rows[1]
for r in rows[2:]:
    print("== kernel:", r[h.index("Kernel Name")][:110])
    for k in KEYS:
        if k in h:
            i = h.index(k)
            print("  %-88s %s %s" % (k, r[i], u[i]))
