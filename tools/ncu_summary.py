"""Print the metrics the roofline / issue analysis needs from an .ncu-rep (ncu -i ... --page raw --csv)."""
import csv, subprocess, sys
WANT = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'dram__throughput.avg.pct_of_peak_sustained_elapsed',
        'smsp__inst_executed.sum', 'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'lts__t_sectors_op_read.sum', 'lts__t_sectors_op_write.sum', 'lts__t_sector_hit_rate.pct', 'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_registers',
        'launch__registers_per_thread', 'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'launch__block_size', 'launch__grid_size', 'smsp__warps_eligible.avg.per_cycle_active',
        'launch__shared_mem_per_block_dynamic', 'smsp__average_warp_latency_issue_stalled', 'smsp__average_warps_issue_stalled']
for rep in sys.argv[1:]:
    out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for vals in rows[2:]:
        d = dict(zip(hdr, vals))
        print('==', rep, d.get('Kernel Name', '')[:90])
        for h, u, v in zip(hdr, units, vals):
            if h in WANT or (h.startswith('smsp__average_warps_issue_stalled') and h.endswith('per_issue_active.ratio') and float(v or 0) > 0.3) \
               or h.startswith('smsp__average_warp_latency_issue_stalled') and float(v or 0) > 1.0:
                print('  %-95s %-14s %s' % (h, u, v))
