"""Key metrics + top-stall SASS lines of one ncu --set full report: python scripts/ncu_report_summary.py <rep> [n_lines]"""
import csv, subprocess, sys
rep=sys.argv[1]
raw=subprocess.run(["ncu","-i",rep,"--page","raw","--csv"],capture_output=True,text=True).stdout
rows=list(csv.reader(raw.splitlines()))
hdr,unit,vals=rows[0],rows[1],rows[2]
want=["gpu__time_duration.sum","sm__cycles_elapsed.avg","sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active","smsp__issue_active.avg.pct_of_peak_sustained_active","sm__inst_executed_pipe_xu","sm__inst_executed_pipe_fma","sm__inst_executed_pipe_alu","sm__pipe_fma_cycles_active","sm__pipe_alu_cycles_active","smsp__inst_executed.sum","dram__bytes_read.sum","dram__bytes_write.sum","l1tex__data_bank_conflicts_pipe_lsu_mem_shared","smsp__average_warps_issue_stalled","sm__warps_active","launch__registers","lts__t_bytes.sum "]
for h,u,v in zip(hdr,unit,vals):
    if any(w in h for w in want) and "no data" not in v and ('.max' not in h and '.min' not in h):
        print(f"{h} [{u}] = {v}")
src=subprocess.run(["ncu","-i",rep,"--page","source","--csv"],capture_output=True,text=True).stdout
rows=list(csv.reader(src.splitlines()))
data=[]
for r in rows[2:]:
    try: data.append((int(r[2]),r))
    except: pass
tot=sum(v for v,_ in data)
print("total samples",tot)
top=sorted(range(len(data)),key=lambda i:-data[i][0])[:int(sys.argv[2]) if len(sys.argv)>2 else 16]
for i in top:
    v,r=data[i]
    print(f"{v:7d} {100*v/tot:5.1f}% exec={r[5]:>9} | {r[1].strip()[:100]}")
    for j in range(max(0,i-2),i):
        print("             ctx:", data[j][1][1].strip()[:100])
