import csv, collections, sys, subprocess
rep=sys.argv[1]
raw=subprocess.run(['ncu','-i',rep,'--page','raw','--csv'],capture_output=True,text=True).stdout
rows=list(csv.reader(raw.splitlines())); d=dict(zip(rows[0],rows[-1]))
for k in ['gpu__time_duration.sum','sm__cycles_elapsed.max','launch__registers_per_thread','smsp__inst_executed.sum','smsp__issue_active.avg.pct_of_peak_sustained_active','sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active','sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active','sm__warps_active.avg.pct_of_peak_sustained_active','l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum']:
    print('%-75s %s'%(k,d.get(k)))
st={k.split('issue_stalled_')[1].split('_per_issue')[0]:float(v) for k,v in d.items() if 'issue_stalled' in k and 'per_issue_active' in k}
print('stalls/issue:', ', '.join('%s %.2f'%(k,v) for k,v in sorted(st.items(), key=lambda x:-x[1])[:9]))
src=subprocess.run(['ncu','-i',rep,'--page','source','--csv'],capture_output=True,text=True).stdout
rows=list(csv.reader(src.splitlines()))
hi=[i for i,r in enumerate(rows) if 'Source' in r and 'Instructions Executed' in r][0]
h=rows[hi]; si=h.index('Source'); ei=h.index('Instructions Executed'); wi=h.index('Warp Stall Sampling (All Samples)')
ops=collections.Counter(); samp=collections.Counter(); tot=0; ts=0; lines=[]
for r in rows[hi+1:]:
    if len(r)<=ei: continue
    try: n=int(r[ei]); s=int(r[wi])
    except: continue
    t=r[si].split()
    op=t[0] if not t[0].startswith('@') else t[1]
    op=op.split('.')[0]
    ops[op]+=n; samp[op]+=s; tot+=n; ts+=s; lines.append((s,n,r[si][:90]))
print("total warp instrs", tot, "samples", ts)
for op,n in ops.most_common(16): print("  %-8s %10d %5.1f%%  samples %5.1f%%"%(op,n,100*n/tot,100*samp[op]/ts))
print("top sampled SASS lines:")
for s,n,l in sorted(lines,reverse=True)[:14]: print("  %5d %9d  %s"%(s,n,l))
