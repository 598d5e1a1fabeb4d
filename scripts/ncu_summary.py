#!/usr/bin/env python
"""Summarise an .ncu-rep (read here, no GPU needed): headline metrics, stall reasons, hottest source lines."""
import csv, io, subprocess, sys, collections
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
H, U, V = rows[0], rows[1], rows[2:]
def get(name):
    for i, h in enumerate(H):
        if h == name:
            return [r[i] for r in V], U[i]
    return None, None
keys = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum", "l1tex__t_bytes.sum", "lts__t_sectors.sum", "l1tex__t_sectors.sum",
        "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "smsp__cycles_active.avg", "sm__cycles_elapsed.avg", "smsp__warps_eligible.avg.per_cycle_active",
        "sm__cycles_active.max", "sm__cycles_active.min", "sm__cycles_active.avg"]
for k in keys:
    v, u = get(k)
    if v: print("%-70s %s %s" % (k, v, u))
print("--- pc-sampling stall reasons (samples) ---")
st = []
for i, h in enumerate(H):
    if h.startswith("smsp__pcsamp_warps_issue_stalled_") and not h.endswith("_not_issued"):
        try: st.append((float(V[0][i]), h.replace("smsp__pcsamp_warps_issue_stalled_", "")))
        except: pass
tot = sum(v for v, _ in st) or 1
for v, h in sorted(st, reverse=True)[:10]:
    print("%6.1f%%  %s" % (100 * v / tot, h))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
# the cuda,sass view is a sequence of per-file blocks: File Path / Function Name / header / per-line rows (+ sass rows)
lines = []
cur_file, Hs = None, None
for r in rows:
    if not r: continue
    if r[0] == "File Path":
        cur_file = r[1].split("/")[-1]; Hs = None; continue
    if r[0] == "Function Name": continue
    if r[0] == "Line No":
        Hs = r; continue
    if Hs is None: continue
    if r[0] in ("-", ""): continue  # sass row
    try:
        si = Hs.index("# Samples"); ii = Hs.index("Instructions Executed")
        lines.append((float(r[si]), float(r[ii]), cur_file, r[0], r[1]))
    except Exception:
        pass
tot = sum(l[0] for l in lines) or 1
toti = sum(l[1] for l in lines) or 1
print("--- hottest source lines: %%samples  %%warp-instructions  file:line  source ---")
for s_, ie, f, ln, text in sorted(lines, reverse=True)[:int(sys.argv[2]) if len(sys.argv) > 2 else 30]:
    print("%5.1f%% %5.1f%%  %s:%s  %s" % (100 * s_ / tot, 100 * ie / toti, f, ln, text.strip()[:100]))
