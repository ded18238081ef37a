"""One line per distinct kernel from an ncu raw CSV (ncu -i X.ncu-rep --page raw --csv)."""
import csv, sys
rows=list(csv.reader(open(sys.argv[1])))
hdr=rows[0]; idx={h:i for i,h in enumerate(hdr)}
want=[('gpu__time_duration.sum','us'),('dram__bytes_read.sum','rdMB'),('dram__bytes_write.sum','wrMB'),
('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed','dram%'),('sm__warps_active.avg.pct_of_peak_sustained_active','occ%'),
('launch__registers_per_thread','regs'),('smsp__issue_active.avg.pct_of_peak_sustained_active','issue%'),
('smsp__cycles_active.avg','cyc_act'),('sm__cycles_elapsed.max','cyc_max'),
('smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio','longsb'),
('smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio','shortsb'),
('smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio','bar'),
('smsp__average_warps_issue_stalled_wait_per_issue_active.ratio','wait'),
('smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio','math'),
('smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio','lg'),
('smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio','mio'),
('smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio','br'),
('smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio','notsel'),
('smsp__average_warps_issue_stalled_membar_per_issue_active.ratio','membar'),
('smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio','disp'),
('smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio','noinst'),
('smsp__thread_inst_executed_per_inst_executed.ratio','thr/inst'),
('smsp__inst_executed.sum','inst'),('lts__t_sectors_op_red.sum','red'),('lts__t_sectors_op_atom.sum','atom'),
('l1tex__data_pipe_lsu_wavefronts_mem_shared.sum','smem_wf')]
seen=set()
for r in rows[2:]:
    name=r[idx['Kernel Name']].split('(')[0].replace('sgs::','').replace('void ','')
    name=name.replace('unsigned long long','u64').replace('unsigned int','u32')[:40]   # keep the template arguments: the u32 and u64 sort passes are different kernels
    if name in seen: continue
    seen.add(name)
    def fmt(v):
        try: return f"{float(v):.4g}"
        except: return v
    print(name, ' '.join(f"{s}={fmt(r[idx[w]])}" for w,s in want if w in idx))
    print()
