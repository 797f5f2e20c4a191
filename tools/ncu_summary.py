#!/usr/bin/env python
"""Summarise an ncu raw-page CSV + source-page CSV (see B200_PROFILING.md)."""
import csv, sys
from collections import Counter
raw, src = sys.argv[1], sys.argv[2]
rows=list(csv.reader(open(raw)))
hdr=rows[0]; units=rows[1]
keys=['gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed','sm__warps_active.avg.pct_of_peak_sustained_active','launch__registers_per_thread','sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active','sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active','l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum','l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum','l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_st.sum','l1tex__data_pipe_lsu_wavefronts_mem_shared.sum','smsp__inst_executed.sum','smsp__issue_active.avg.pct_of_peak_sustained_active','sm__cycles_elapsed.avg','sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_adu.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active','smsp__warps_eligible.avg.per_cycle_active','l1tex__throughput.avg.pct_of_peak_sustained_active','lts__throughput.avg.pct_of_peak_sustained_elapsed']
for k in keys:
    if k in hdr:
        i=hdr.index(k); print(k, units[i], [r[i] for r in rows[2:]])
print('-- stalls per issue')
for h in hdr:
    if 'average_warps_issue_stalled' in h and 'per_issue_active' in h:
        i=hdr.index(h); print('  ',h.replace('smsp__average_warps_issue_stalled_','').replace('_per_issue_active.ratio',''), [r[i] for r in rows[2:]])
rows=list(csv.reader(open(src)))
hdr_idx=[i for i,r in enumerate(rows) if r and r[0]=='Address']
h=rows[hdr_idx[0]]
end=hdr_idx[1]-1 if len(hdr_idx)>1 else len(rows)
body=rows[hdr_idx[0]+1:end]
si=h.index('# Samples'); ii=h.index('Instructions Executed')
tot=sum(int(r[si]) for r in body if len(r)>si and r[si].isdigit())
print('total samples',tot,'instrs',len(body))
c=Counter(); ci=Counter()
for r in body:
    if len(r)<=si or not r[si].isdigit(): continue
    t=r[1].split()
    op=t[0] if not t[0].startswith('@') else t[1]
    c[op]+=int(r[si]); ci[op]+=int(r[ii])
for op,v in c.most_common(16): print('  ',op, v, round(100*v/tot,1), ci[op])
top=sorted(body,key=lambda r:-int(r[si]) if len(r)>si and r[si].isdigit() else 0)[:16]
for r in top: print('  ',r[si], r[ii], r[1][:90])
