import csv,sys
rows=list(csv.reader(open(sys.argv[1])))
hdr=rows[0]; units=rows[1]
keys=['Kernel Name','gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','dram__throughput.avg.pct_of_peak_sustained_elapsed','l1tex__throughput.avg.pct_of_peak_sustained_active','lts__throughput.avg.pct_of_peak_sustained_elapsed','sm__throughput.avg.pct_of_peak_sustained_elapsed','l1tex__t_sector_hit_rate.pct','lts__t_sector_hit_rate.pct','sm__warps_active.avg.pct_of_peak_sustained_active','launch__registers_per_thread','launch__occupancy_limit_registers','launch__occupancy_limit_shared_mem','smsp__inst_executed.sum','lts__t_sectors_srcunit_tex_op_read.sum','sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active','smsp__issue_active.avg.pct_of_peak_sustained_active','sm__cycles_elapsed.avg','l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed','l1tex__data_pipe_lsu_wavefronts_mem_shared.sum','l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum','smsp__average_warp_latency_per_inst_issued.ratio','launch__shared_mem_per_block_dynamic']
for r in rows[2:]:
    print('-----')
    for k in keys:
        if k in hdr:
            i=hdr.index(k); print(k, r[i], units[i])
    for i,h in enumerate(hdr):
        if 'issue_stalled' in h and h.endswith('_per_issue_active.ratio') and float(r[i])>0.1: print('  ',h.replace('smsp__average_warps_issue_stalled_','').replace('_per_issue_active.ratio',''), r[i])
