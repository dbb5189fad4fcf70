import csv,sys,subprocess
rep=sys.argv[1]
raw=subprocess.run(['ncu','-i',rep,'--page','raw','--csv'],capture_output=True,text=True).stdout
rows=list(csv.reader(raw.splitlines()))
hdr=rows[0]; units=rows[1]
want=['Kernel Name','gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active','sm__throughput.avg.pct_of_peak_sustained_elapsed','lts__throughput.avg.pct_of_peak_sustained_elapsed','l1tex__throughput.avg.pct_of_peak_sustained_elapsed','gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed','sm__cycles_elapsed.max','smsp__inst_executed.sum','launch__grid_size','launch__registers_per_thread','launch__occupancy_limit_shared_mem']
for w in want:
    for i,h in enumerate(hdr):
        if h==w: print('%-70s %-12s %s'%(h,units[i],[r[i][:50] for r in rows[2:]]))
src=subprocess.run(['ncu','-i',rep,'--page','source','--csv','--print-source','sass'],capture_output=True,text=True).stdout
rows=list(csv.reader(src.splitlines()))
secs=[i for i,r in enumerate(rows) if r and r[0]=='Kernel Name']
seen=set()
for si,s in enumerate(secs):
    name=rows[s][1]
    if name in seen: continue
    seen.add(name)
    hdr=rows[s+1]
    end=secs[si+1] if si+1<len(secs) else len(rows)
    body=rows[s+2:end]
    ci=hdr.index('# Samples'); srcc=hdr.index('Source'); ie=hdr.index('Instructions Executed')
    stall_cols=[(j,h) for j,h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
    tot=sum(int(r[ci]) for r in body if len(r)>ci and r[ci].isdigit())
    print('==== ',name[:70],'total samples',tot)
    agg={}
    for r in body:
        if len(r)>ci and r[ci].isdigit():
            for j,h in stall_cols:
                if r[j].isdigit(): agg[h[6:]]=agg.get(h[6:],0)+int(r[j])
    print({k:v for k,v in sorted(agg.items(), key=lambda kv:-kv[1])[:8]})
    top=sorted([r for r in body if len(r)>ci and r[ci].isdigit()], key=lambda r:-int(r[ci]))[:int(sys.argv[2]) if len(sys.argv)>2 else 16]
    for r in top:
        st={h[6:]:int(r[j]) for j,h in stall_cols if r[j].isdigit() and int(r[j])>0}
        st=dict(sorted(st.items(), key=lambda kv:-kv[1])[:2])
        print('%6s %5.1f%% ie=%8s %-64s %s'%(r[ci],100*int(r[ci])/tot,r[ie],r[srcc].strip()[:64],st))
