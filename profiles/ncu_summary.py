import csv, subprocess, sys
rep=sys.argv[1]
raw=subprocess.run(['ncu','-i',rep,'--page','raw','--csv','--print-units','base'],capture_output=True,text=True).stdout
rows=list(csv.reader(raw.splitlines()))
hdr=rows[0]; idx={h:i for i,h in enumerate(hdr)}
want=['gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed','sm__warps_active.avg.pct_of_peak_sustained_active','launch__registers_per_thread','sm__throughput.avg.pct_of_peak_sustained_elapsed','lts__t_sector_hit_rate.pct','l1tex__t_sector_hit_rate.pct','lts__throughput.avg.pct_of_peak_sustained_elapsed','smsp__issue_active.avg.pct_of_peak_sustained_active','smsp__inst_executed.sum','l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed','l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum','sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active','launch__grid_size','launch__occupancy_limit_registers','launch__occupancy_limit_shared_mem']
stalls=[h for h in hdr if h.startswith('smsp__average_warps_issue_stalled') and h.endswith('per_issue_active.ratio')]
for r in rows[2:]:
    print('====', r[idx['Kernel Name']][:50])
    for w in want:
        if w in idx: print('  ',w,'=',r[idx[w]])
    st=sorted([(float(r[idx[h]]),h.replace('smsp__average_warps_issue_stalled_','').replace('_per_issue_active.ratio','')) for h in stalls],reverse=True)[:6]
    print('   stalls:', ', '.join(f'{n}={v:.2f}' for v,n in st))
