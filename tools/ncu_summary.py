import csv, subprocess, sys
rep=sys.argv[1]
raw=subprocess.run(['ncu','-i',rep,'--page','raw','--csv'],capture_output=True,text=True).stdout
rows=list(csv.reader(raw.splitlines()))
h=rows[0]
keys=['Kernel Name','gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed','sm__warps_active.avg.pct_of_peak_sustained_active','launch__registers_per_thread','launch__grid_size','launch__block_size','smsp__inst_executed.sum','smsp__issue_active.avg.pct_of_peak_sustained_active','l1tex__data_pipe_lsu_wavefronts_mem_shared.sum','sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active','lts__t_sector_hit_rate.pct']
for r in rows[2:]:
    for k in keys:
        if k in h: print(k.split('.')[0][:40], '=', r[h.index(k)][:90])
src=subprocess.run(['ncu','-i',rep,'--page','source','--csv'],capture_output=True,text=True).stdout
rows=list(csv.reader(src.splitlines()))
hdr=rows[1]; idx={k:i for i,k in enumerate(hdr)}
sc=[k for k in hdr if k.startswith('stall_') and 'Not Issued' not in k]
tot={k:0 for k in sc}; lines=[]
for r in rows[2:]:
    if len(r)<len(hdr): continue
    try: n=int(r[idx['# Samples']])
    except: continue
    for k in sc: tot[k]+=int(r[idx[k]] or 0)
    lines.append((n,r[idx['Source']].strip(),{k[6:]:int(r[idx[k]] or 0) for k in sc if int(r[idx[k]] or 0)>0}))
s=sum(tot.values()) or 1
print({k[6:]:round(100*v/s,1) for k,v in tot.items() if v>0.02*s})
lines.sort(key=lambda x:-x[0])
for n,srcl,st in lines[:int(sys.argv[2]) if len(sys.argv)>2 else 12]: print(n,srcl[:60],st)
