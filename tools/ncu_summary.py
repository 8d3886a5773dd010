"""Summarise an .ncu-rep: key counters, stall mix, top stall instructions.  usage: python tools/ncu_summary.py report.ncu-rep"""
import csv, sys, subprocess
rep=sys.argv[1]
raw=subprocess.run(['ncu','-i',rep,'--page','raw','--csv'],capture_output=True,text=True).stdout
rows=list(csv.reader(raw.splitlines()))
hdr=rows[0]; units=rows[1]; vals=rows[2]
want=['gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed','l1tex__throughput.avg.pct_of_peak_sustained_elapsed','lts__throughput.avg.pct_of_peak_sustained_elapsed','sm__warps_active.avg.pct_of_peak_sustained_active','launch__registers_per_thread','launch__grid_size','sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active','smsp__inst_executed.sum','l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum','l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum','l1tex__t_sector_hit_rate.pct','lts__t_sector_hit_rate.pct','sm__cycles_elapsed.avg','smsp__issue_active.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active','smsp__thread_inst_executed_per_inst_executed.ratio','l1tex__t_output_wavefronts_pipe_lsu_mem_global_op_ld.sum','l1tex__data_pipe_lsu_wavefronts_mem_shared.sum','smsp__warps_eligible.avg.per_cycle_active','l1tex__lsu_writeback_active.avg.pct_of_peak_sustained_elapsed','l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum','sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active','l1tex__f_wavefronts.avg.pct_of_peak_sustained_elapsed' ]
for i,h in enumerate(hdr):
    if h in want: print(h, units[i], vals[i])
st=[]
for i,h in enumerate(hdr):
    if 'issue_stalled' in h and h.endswith('per_issue_active.ratio'):
        try: st.append((float(vals[i]),h.replace('smsp__average_warps_issue_stalled_','').replace('_per_issue_active.ratio','')))
        except: pass
print('stalls:', ', '.join('%s=%.2f'%(h,v) for v,h in sorted(st,reverse=True)[:7]))
src=subprocess.run(['ncu','-i',rep,'--page','source','--csv'],capture_output=True,text=True).stdout
rows=list(csv.reader(src.splitlines()))[2:]
tot=sum(int(r[2]) for r in rows)
print('top stall instructions (of %d samples):'%tot)
for i,r in enumerate(rows):
    s=int(r[2])
    if s>tot*0.02: print('  %4d %5.1f%%  %s | execs %s'%(i,100*s/tot,r[1].strip(),r[5]))
