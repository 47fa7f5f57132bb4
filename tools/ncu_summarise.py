#!/usr/bin/env python
"""Summary of an ncu --set full capture for the tracked profiles: python tools/ncu_summarise.py X.ncu-rep <workload> [out.json]
Writes / updates profiles/ncu_summary.json[workload][kernel] (bench.py reads `traffic` and the texture-pipe figure from there)
and prints a table. One entry per kernel name: the LAST launch of that kernel in the capture."""
import csv, io, json, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rep, workload = sys.argv[1], sys.argv[2]
out_path = sys.argv[3] if len(sys.argv) > 3 else os.path.join(ROOT, "profiles", "ncu_summary.json")
txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
h, units = rows[0], rows[1]
col = {k: i for i, k in enumerate(h)}
SCALE = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "ms": 1e-3, "us": 1e-6, "ns": 1e-9, "s": 1.0, "msecond": 1e-3, "usecond": 1e-6, "nsecond": 1e-9, "second": 1.0}


def val(r, k):
    if k not in col or r[col[k]] in ("", "n/a"):
        return None
    v = float(r[col[k]].replace(",", ""))
    return v * SCALE.get(units[col[k]], 1.0)


WANT = {"duration_us": ("gpu__time_duration.sum", 1e6), "registers": ("launch__registers_per_thread", 1), "dram_read_bytes": ("dram__bytes_read.sum", 1),
        "dram_write_bytes": ("dram__bytes_write.sum", 1), "warps_active_pct": ("sm__warps_active.avg.pct_of_peak_sustained_active", 1),
        "issue_active_pct": ("smsp__issue_active.avg.pct_of_peak_sustained_active", 1), "threads_per_inst": ("smsp__thread_inst_executed_per_inst_executed.ratio", 1),
        "warp_instructions": ("smsp__inst_executed.sum", 1), "l1tex_hit_pct": ("l1tex__t_sector_hit_rate.pct", 1), "l2_hit_pct": ("lts__t_sector_hit_rate.pct", 1),
        "tex_data_pipe_pct": ("l1tex__data_pipe_tex_wavefronts.avg.pct_of_peak_sustained_elapsed", 1), "dram_pct_of_peak": ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", 1),
        "sm_throughput_pct": ("sm__throughput.avg.pct_of_peak_sustained_elapsed", 1), "alu_pipe_pct": ("sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", 1),
        "fma_pipe_pct": ("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", 1)}
summary = {}
for r in rows[2:]:
    name = re.sub(r"\(.*", "", r[col["Kernel Name"]]).split("::")[-1].strip()
    name = re.sub(r"<.*", "", name)
    e = {}
    for k, (m, s) in WANT.items():
        v = val(r, m)
        if v is not None:
            e[k] = round(v * s, 3)
    if "dram_read_bytes" in e and "dram_write_bytes" in e:
        e["dram_bytes"] = e["dram_read_bytes"] + e["dram_write_bytes"]
    summary[name] = e
allsum = {}
if os.path.exists(out_path):
    allsum = json.load(open(out_path))
allsum.setdefault(workload, {}).update(summary)
allsum[workload]["_source"] = os.path.basename(rep) + " (ncu --set full --clock-control none; cold-cache, serialised launches)"
json.dump(allsum, open(out_path, "w"), indent=1, sort_keys=True)
for k, e in summary.items():
    print(f"{k:22s} {e.get('duration_us', 0):8.1f} us  regs {e.get('registers', 0):3.0f}  warps {e.get('warps_active_pct', 0):5.1f}%  issue {e.get('issue_active_pct', 0):5.1f}%  "
          f"thr/inst {e.get('threads_per_inst', 0):4.1f}  L1 {e.get('l1tex_hit_pct', 0):5.1f}%  L2 {e.get('l2_hit_pct', 0):5.1f}%  tex pipe {e.get('tex_data_pipe_pct', 0):5.1f}%  "
          f"DRAM {e.get('dram_bytes', 0) / 1e6:8.1f} MB ({e.get('dram_pct_of_peak', 0):4.1f}%)")
