"""Turn one `ncu --set full` report of k_trace into what profiles/ carries:
    profiles/<tag>_k_trace_raw.csv           the raw page (one launch)
    profiles/<tag>_k_trace_source_lines.csv  the hottest source lines (share of executed warp instructions, of stall samples, threads per instruction)
    profiles/r2_summary.json                 the few counters bench.py quotes in its roofline block, keyed by workload and tree size
usage: python tools/ncu_summary.py <report.ncu-rep> <tag> <workload> <rays> <wide_nodes>"""
import collections
import csv
import io
import json
import os
import re
import subprocess
import sys

rep, tag, workload, rays, wide_nodes = sys.argv[1], sys.argv[2], sys.argv[3], int(sys.argv[4]), int(sys.argv[5])
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
open(os.path.join(ROOT, "profiles", f"{tag}_k_trace_raw.csv"), "w").write(raw)
rows = list(csv.reader(io.StringIO(raw)))
hdr, vals = rows[0], rows[2]
m = dict(zip(hdr, vals))


def f(name):
    return float(m[name].replace(",", ""))


def unit(name):
    return rows[1][hdr.index(name)]


def to_bytes(name):
    v, u = f(name), unit(name)
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}[u]


dur = f("gpu__time_duration.sum")
dur_ms = dur * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}[unit("gpu__time_duration.sum")]
inst = f("smsp__inst_executed.sum") if "smsp__inst_executed.sum" in m else f("sm__inst_executed.sum")
entry = {
    "source": f"ncu --set full, profiles/{tag}_k_trace_raw.csv ({workload}, {rays} rays, one launch)",
    "wide_nodes": wide_nodes, "rays": rays, "duration_ms": dur_ms,
    "dram_bytes": to_bytes("dram__bytes_read.sum") + to_bytes("dram__bytes_write.sum"),
    "dram_read_bytes": to_bytes("dram__bytes_read.sum"), "dram_write_bytes": to_bytes("dram__bytes_write.sum"),
    "lts_sectors": f("lts__t_sectors.sum"), "l2_hit_rate_pct": f("lts__t_sector_hit_rate.pct"),
    "l2_throughput_pct": f("lts__throughput.avg.pct_of_peak_sustained_elapsed") if "lts__throughput.avg.pct_of_peak_sustained_elapsed" in m else None,
    "issue_active_pct": f("smsp__issue_active.avg.pct_of_peak_sustained_active"),
    "warp_instructions": inst, "warp_instructions_per_ray": inst / rays,
    "threads_per_instruction": f("smsp__thread_inst_executed_per_inst_executed.ratio"),
    "registers": f("launch__registers_per_thread"), "pipe_alu_pct": f("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"),
    "pipe_fma_pct": f("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active"),
    "l1_data_pipe_pct": f("l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed"),
    "warps_active_pct": f("sm__warps_active.avg.pct_of_peak_sustained_active"),
}
path = os.path.join(ROOT, "profiles", "r2_summary.json")
allv = json.load(open(path)) if os.path.exists(path) else {}
allv[workload] = entry
json.dump(allv, open(path, "w"), indent=1)
print(json.dumps(entry, indent=1))

src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
lines = list(csv.reader(io.StringIO(src)))
h, fname = None, ""
agg = collections.OrderedDict()
for r in lines:
    if len(r) >= 2 and r[0] == "File Path":
        fname = os.path.basename(r[1])
        continue
    if len(r) > 4 and r[0] == "Line No":
        h = r
        continue
    if not h or len(r) != len(h) or not r[0].strip().isdigit():
        continue
    try:
        ie = float(r[h.index("Instructions Executed")]); te = float(r[h.index("Thread Instructions Executed")]); sm = float(r[h.index("# Samples")] or 0)
    except ValueError:
        continue
    if ie <= 0:
        continue
    a = agg.setdefault((fname, int(r[0]), r[1].strip()), [0.0, 0.0, 0.0])
    a[0] += ie; a[1] += te; a[2] += sm
tot_i = sum(a[0] for a in agg.values()) or 1.0
tot_s = sum(a[2] for a in agg.values()) or 1.0
top = sorted(agg.items(), key=lambda kv: -kv[1][0])[:80]
with open(os.path.join(ROOT, "profiles", f"{tag}_k_trace_source_lines.csv"), "w") as fo:
    w = csv.writer(fo)
    w.writerow(["file", "line", "inst_executed_pct", "samples_pct", "avg_threads", "source"])
    for (fn, ln, text), a in top:
        w.writerow([fn, ln, "%.2f" % (100 * a[0] / tot_i), "%.2f" % (100 * a[2] / tot_s), "%.1f" % (a[1] / a[0] if a[0] else 0), text[:100]])
print("source lines:", len(agg))
