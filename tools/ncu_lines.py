#!/usr/bin/env python
"""Per-source-line aggregation of an ncu capture: tools/ncu_lines.py <rep.ncu-rep> <object.o> <mangled kernel name> [units]
Disassembles the object's cubin with line info (nvdisasm -g) and joins it with `ncu --page source --csv`."""
import collections, csv, os, re, subprocess, sys, tempfile
rep, obj, fn = sys.argv[1:4]
units = float(sys.argv[4]) if len(sys.argv) > 4 else 1.0
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, check=True, capture_output=True)
cub = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
sass = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cub)], capture_output=True, text=True).stdout.split("\n")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
start = None
for i, l in enumerate(sass):
    if l.strip().startswith(".section") and (".text." + fn) in l:
        start = i; break
assert start is not None, "kernel not found in the object"
cur = ("?", 0); instrs = []
for l in sass[start + 1:]:
    if l.strip().startswith(".section"): break
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m: cur = (m.group(1).split("/")[-1], int(m.group(2))); continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
    if m: instrs.append((int(m.group(1), 16), cur, m.group(2)))
rows = list(csv.reader(src.split("\n")))
hdr = rows[1]; data = [r for r in rows[2:] if r and r[0] != "Kernel Name" and len(r) == len(hdr)]
ia, ie, ismp, isrc = hdr.index("Address"), hdr.index("Instructions Executed"), hdr.index("# Samples"), hdr.index("Source")
base = int(data[0][ia], 16)
byoff = {o: (c, t) for o, c, t in instrs}
agg = collections.Counter(); smp = collections.Counter(); op = collections.Counter(); tot = tots = 0
for r in data:
    off = int(r[ia], 16) - base; n = int(r[ie]); s_ = int(r[ismp])
    c, t = byoff.get(off, (("?", 0), ""))
    agg[c] += n; smp[c] += s_; tot += n; tots += s_
    o = r[isrc].split()[0] if not r[isrc].strip().startswith("@") else r[isrc].split()[1]
    op[o.split(".")[0]] += n
print(f"total warp instructions {tot}  per unit {tot / units:.1f}  samples {tots}")
rr = list(csv.reader(raw.split("\n")))
if len(rr) > 2:
    h, v = rr[0], rr[2]
    for key in ("gpu__time_duration.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
                "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
                "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
                "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "sass__inst_executed_local_loads",
                "dram__bytes_read.sum", "dram__bytes_write.sum", "smsp__average_warp_latency_per_inst_issued.ratio",
                "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
                "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
                "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio", "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
                "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
                "smsp__thread_inst_executed_per_inst_executed.ratio"):
        if key in h: print(f"  {key:95s} {v[h.index(key)]}")
print("--- by opcode")
for k, v in op.most_common(16): print(f"{k:10s} {v / tot:6.3f}")
print("--- by source line (share of executed warp instructions, share of stall samples)")
for (f, l), v in agg.most_common(int(os.environ.get("TOP", "50"))):
    print(f"{f}:{l:4d} inst={v / tot:6.3f} ({v / units:8.1f}/unit) samples={smp[(f, l)] / max(tots, 1):6.3f}")
