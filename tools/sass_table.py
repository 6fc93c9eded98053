#!/usr/bin/env python
"""Per-kernel SASS counts (instructions, DMMA, LDGSTS = cp.async, UBLKCP / UTMALDG = bulk / tensor TMA copies) of the shipped
library, and the inner loop of the N > 32 Hessian kernel.   python tools/sass_table.py > profiles/r02_sass_tensor_core_kernels.txt"""
import os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "mpc-ntm-control_b200", "lib", "libntm_mpc.so")
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
fn, rows, body = None, {}, {}
for l in sass.split("\n"):
    m = re.search(r"Function : (\S+)", l)
    if m: fn = m.group(1); rows[fn] = [0, 0, 0, 0]; body[fn] = []; continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(.*?);", l)
    if fn and m:
        t = m.group(1); rows[fn][0] += 1; body[fn].append(t.strip())
        if "DMMA" in t: rows[fn][1] += 1
        if "LDGSTS" in t: rows[fn][2] += 1
        if "UBLKCP" in t or "UTMALDG" in t: rows[fn][3] += 1
for f in sorted(rows):
    r = rows[f]
    if r[1] or r[2] or r[3]:
        print(f"{f:90s} instr {r[0]:6d}  DMMA {r[1]:4d}  LDGSTS {r[2]:3d}  UBLKCP/UTMALDG {r[3]}")
f = next((k for k in rows if "hessian_grad_dmma_kernelILi4" in k), None)
if f:
    b = body[f]
    i0 = next(i for i, t in enumerate(b) if "LDS.128" in t and any("DMMA" in u for u in b[i:i + 12]))
    print(f"\n--- {f}: first super-tile loop (one LDS.128 per fragment feeds two DMMA.8x8x4; 8 rows of K per trip)")
    for t in b[i0 - 2:i0 + 44]: print("    " + t)
