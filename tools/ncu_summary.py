#!/usr/bin/env python
"""Summarise an ncu raw-page CSV + source-page CSV (see B200_PROFILING.md).
    python tools/ncu_summary.py raw.csv source.csv [--dram-json out.json n_qubits kernel_regex]
--dram-json writes the per-launch DRAM bytes of the first launch whose kernel name matches (bench.py reads that file for
`roofline.traffic`)."""
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

if '--dram-json' in sys.argv:
    import json, re
    j=sys.argv.index('--dram-json'); out,nq,rx=sys.argv[j+1],int(sys.argv[j+2]),sys.argv[j+3]
    rows=list(csv.reader(open(raw))); hdr=rows[0]; units=rows[1]
    kn=hdr.index('Kernel Name')
    def val(r,key):
        i=hdr.index(key); v=float(r[i].replace(',','')); u=units[i].lower()
        return v*{'gbyte':1e9,'mbyte':1e6,'kbyte':1e3,'byte':1.0}.get(u,1.0)
    for r in rows[2:]:
        if re.search(rx,r[kn]):
            rec={'n_qubits':nq,'kernel':r[kn][:120],'dram_bytes_read':val(r,'dram__bytes_read.sum'),
                 'dram_bytes_write':val(r,'dram__bytes_write.sum'),'gpu_time_ms':val(r,'gpu__time_duration.sum')/1e6 if units[hdr.index('gpu__time_duration.sum')].lower().startswith('n') else None,
                 'source':'ncu --set full --clock-control none, raw page '+raw}
            json.dump(rec,open(out,'w'),indent=1); print('wrote',out,rec); break
