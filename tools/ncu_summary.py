"""Key metrics and top warp-stall reasons of every kernel in an .ncu-rep (ncu --set full capture): python tools/ncu_summary.py <report>."""
import csv,sys,subprocess
rep=sys.argv[1]
out=subprocess.run(["ncu","-i",rep,"--page","raw","--csv"],capture_output=True,text=True).stdout
rows=list(csv.reader(out.splitlines()))
hdr=rows[0]; units=rows[1]
want=["Kernel Name","gpu__time_duration.sum","dram__bytes_read.sum","dram__bytes_write.sum","gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed","sm__warps_active.avg.pct_of_peak_sustained_active","launch__registers_per_thread","launch__grid_size","launch__block_size","sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active","smsp__issue_active.avg.pct_of_peak_sustained_active","l1tex__t_sector_hit_rate.pct","lts__t_sector_hit_rate.pct","smsp__inst_executed.sum","sm__cycles_elapsed.avg.per_second"]
for r in rows[2:]:
    d=dict(zip(hdr,r))
    for w in want:
        if w in d: print(w,"=",d[w],units[hdr.index(w)])
    st=[(k,float(v.replace(',',''))) for k,v in d.items() if 'warps_issue_stalled' in k and k.endswith('per_issue_active.ratio') and v not in ('','n/a')]
    st.sort(key=lambda kv:-kv[1])
    print("stalls per issue:", ", ".join("%s %.2f"%(k.replace('smsp__average_warps_issue_stalled_','').replace('_per_issue_active.ratio',''),v) for k,v in st[:7]))
    print()
