"""Print a compact table of the metrics that matter for the list-walking sweeps from .ncu-rep files."""
import csv, subprocess, sys, io
KEYS = ['gpu__time_duration.sum','sm__cycles_elapsed.max','smsp__inst_executed.sum','l1tex__data_pipe_lsu_wavefronts.sum',
 'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed','l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum',
 'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum','l1tex__t_sector_hit_rate.pct','lts__t_sector_hit_rate.pct',
 'dram__bytes_read.sum','dram__bytes_write.sum','gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
 'smsp__issue_active.avg.pct_of_peak_sustained_active','sm__warps_active.avg.pct_of_peak_sustained_active',
 'launch__registers_per_thread','smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
 'smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio','smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
 'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio','smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio',
 'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio','smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio',
 'smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio','smsp__average_warps_issue_stalled_tex_throttle_per_issue_active.ratio',
 'smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio','smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio',
 'smsp__thread_inst_executed_per_inst_executed.ratio','l1tex__lsu_writeback_active.avg.pct_of_peak_sustained_elapsed',
 'l1tex__m_xbar2l1tex_read_sectors.sum','sm__inst_executed_pipe_lsu.sum','launch__grid_size']
for rep in sys.argv[1:]:
    out = subprocess.run(['ncu','-i',rep,'--page','raw','--csv'],capture_output=True,text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    print('==', rep)
    names = [r[hdr.index('Kernel Name')].split('(')[0][-28:] for r in rows[2:]]
    print('%-90s'%'metric', names)
    for k in KEYS:
        if k in hdr:
            i = hdr.index(k)
            print('%-90s'%(k[-88:]+' '+units[i]), [r[i][:12] for r in rows[2:]])
