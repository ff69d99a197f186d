"""Turn the ncu artefacts of a gpurun call into the small text summaries kept under profiles/.

    python tools/ncu_summarize.py <tag> <launches.csv> <full.ncu-rep> "<command line that was profiled>"

Writes profiles/<tag>_launches_summary.csv (per kernel: launches, total, average, share of the
step), profiles/<tag>_ncu_full_key_metrics.csv (first captured launch of every kernel) and
refreshes profiles/ncu_traffic.json (DRAM bytes per launch, read by bench.py's roofline block).
Runs here on the CPU box: `ncu -i ... --page raw --csv` needs no GPU."""
import csv
import io
import json
import os
import re
import subprocess
import sys
from collections import OrderedDict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread",
        "launch__waves_per_multiprocessor", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "sm__cycles_elapsed.max"]


def short(name):
    name = re.sub(r"\(.*", "", name).strip()
    name = name.replace("void ", "").replace("dca::", "").replace("(anonymous namespace)::", "")
    return re.sub(r"^.*unnamed>::", "", name)


def launches(path, tag, cmd):
    rows = [r for r in csv.reader(l for l in open(path) if not l.startswith("=="))]
    head = rows[0]
    ki, mi, vi = head.index("Kernel Name"), head.index("Metric Name"), head.index("Metric Value")
    ui = head.index("Metric Unit")
    agg = OrderedDict()
    for r in rows[1:]:
        if len(r) <= vi or r[mi] != "gpu__time_duration.sum":
            continue
        v = float(r[vi].replace(",", ""))
        v = v / 1e3 if r[ui] in ("ns", "nsecond") else v
        a = agg.setdefault(short(r[ki]), [0, 0.0])
        a[0] += 1
        a[1] += v
    total = sum(a[1] for a in agg.values())
    out = os.path.join(ROOT, "profiles", "%s_launches_summary.csv" % tag)
    with open(out, "w") as f:
        f.write("# ncu --metrics gpu__time_duration.sum --clock-control none, %s\n" % cmd)
        f.write("# per-launch times are cold-cache and serialised: compare SHARES\n")
        f.write("kernel,launches,total_us,avg_us,share\n")
        for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write("%s,%d,%.1f,%.2f,%.4f\n" % (k, n, t, t / n, t / total))
    return out


def full(path, tag, cmd):
    txt = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    head, units = rows[0], rows[1]
    ki = head.index("Kernel Name")
    seen = OrderedDict()
    for r in rows[2:]:
        seen.setdefault(short(r[ki]), r)
    out = os.path.join(ROOT, "profiles", "%s_ncu_full_key_metrics.csv" % tag)
    traffic = {"source": "ncu --set full --clock-control none, %s, first captured launch per kernel (profiles/%s)"
                         % (cmd, os.path.basename(out)), "kernels": {}}
    with open(out, "w") as f:
        f.write("# ncu --set full --clock-control none --import-source on, %s; first captured launch of each kernel\n" % cmd)
        for k, r in seen.items():
            f.write("---- %s\n" % k)
            vals = {}
            for key in KEYS:
                if key in head:
                    i = head.index(key)
                    f.write("%s,%s,%s\n" % (key, r[i], units[i]))
                    vals[key] = (r[i], units[i])

            def to_bytes(key):
                v, u = vals.get(key, ("0", "byte"))
                v = float(v.replace(",", "")) if v else 0.0
                return v * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}.get(u, 1.0)
            rd, wr = to_bytes("dram__bytes_read.sum"), to_bytes("dram__bytes_write.sum")
            t, tu = vals.get("gpu__time_duration.sum", ("0", "us"))
            t = float(t.replace(",", "")) * {"ms": 1e3, "us": 1.0, "ns": 1e-3, "msecond": 1e3, "usecond": 1.0,
                                              "nsecond": 1e-3}.get(tu, 1.0)
            traffic["kernels"][k] = {"dram_bytes_read": rd, "dram_bytes_write": wr, "dram_bytes": rd + wr, "gpu_time_us": t}
    if not os.environ.get("NCU_KEEP_TRAFFIC"):      # captures of other kernels must not replace the plmDCA table
        json.dump(traffic, open(os.path.join(ROOT, "profiles", "ncu_traffic.json"), "w"), indent=1)
    return out


if __name__ == "__main__":
    tag, lpath, fpath, cmd = sys.argv[1:5]
    if lpath != "-":
        print(launches(lpath, tag, cmd))
    if fpath != "-":
        print(full(fpath, tag, cmd))
