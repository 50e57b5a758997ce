# summarise an .ncu-rep: one block per captured launch with the metrics the roofline discussion needs
import csv, subprocess, sys
rep = sys.argv[1]
raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}
want = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'dram__cycles_active.avg.pct_of_peak_sustained_elapsed', 'lts__t_bytes.sum', 'lts__t_sector_hit_rate.pct', 'l1tex__t_sector_hit_rate.pct',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size',
        'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem', 'smsp__inst_executed.sum',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio', 'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio']
for r in rows[2:]:
    print('==', r[idx['Kernel Name']][:110], ' grid', r[idx.get('Grid Size', 0)], 'block', r[idx.get('Block Size', 0)])
    for w in want:
        if w in idx:
            print('   %-88s %14s %s' % (w, r[idx[w]], units[idx[w]]))
