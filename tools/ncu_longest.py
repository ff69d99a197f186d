"""Key metrics of an ncu report, one row per (kernel, grid): the longest captured launch.

    python tools/ncu_longest.py <report.ncu-rep> <out.csv> "<note: the command that was profiled>"
"""
import csv, subprocess, sys, re, io
rep, out, note = sys.argv[1], sys.argv[2], sys.argv[3]
txt = subprocess.run(["ncu","-i",rep,"--page","raw","--csv"],capture_output=True,text=True).stdout
rows=list(csv.reader(io.StringIO(txt)))
hdr,units=rows[0],rows[1]
KEYS=["gpu__time_duration.sum","launch__grid_size","launch__block_size","launch__registers_per_thread","launch__waves_per_multiprocessor",
"dram__bytes_read.sum","dram__bytes_write.sum","sm__warps_active.avg.pct_of_peak_sustained_active","smsp__inst_executed.sum",
"smsp__issue_active.avg.pct_of_peak_sustained_active","sm__inst_executed_pipe_tensor.avg.pct_of_peak_sustained_active",
"sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active","sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
"l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed","lts__throughput.avg.pct_of_peak_sustained_elapsed",
"gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
"smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio","smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
"smsp__average_warps_issue_stalled_wait_per_issue_active.ratio","smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
"smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio","smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio"]
ki=hdr.index("Kernel Name"); gi=hdr.index("Grid Size") if "Grid Size" in hdr else None
idx={k:hdr.index(k) for k in KEYS if k in hdr}
best={}
def short(n):
    n=re.sub(r"\(.*","",n).replace("void ","").replace("dca::","")
    return n
for r in rows[2:]:
    key=(short(r[ki]), r[gi] if gi is not None else "")
    t=float(r[idx["gpu__time_duration.sum"]].replace(",",""))
    if key not in best or t>best[key][0]: best[key]=(t,r)
with open(out,"w") as f:
    f.write("# "+note+"\n# one row per (kernel, grid): the longest captured launch\n")
    f.write("kernel,grid,"+",".join(k for k in KEYS if k in idx)+"\n")
    f.write("unit,,"+",".join(units[idx[k]] for k in KEYS if k in idx)+"\n")
    for (k,g),(t,r) in best.items():
        f.write('"%s","%s",'%(k,g)+",".join(r[idx[kk]].replace(",","") for kk in KEYS if kk in idx)+"\n")
print(open(out).read()[:6000])
