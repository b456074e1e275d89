#!/usr/bin/env python3
"""Static check of a built library's hot loops (dev tool): for every kernel whose name contains argv[2], list the loops
with their instruction count and the sum of the stall fields ptxas put on them (the issue time of one iteration of a lone
warp).  usage: tools/sass_loops.py lib.so k_step_ws"""
import re
import subprocess
import sys

out = subprocess.run(["cuobjdump", "-sass", sys.argv[1]], capture_output=True, text=True).stdout.split("\n")
funcs, cur, i = {}, None, 0
while i < len(out):
    m = re.match(r"\s+Function : (\S+)", out[i])
    if m:
        cur = m.group(1); funcs[cur] = []; i += 1; continue
    m = re.match(r"\s+/\*([0-9a-f]{4,6})\*/\s+(.*?);\s*/\* (0x[0-9a-f]+) \*/", out[i])
    if m and cur:
        hi = 0
        if i + 1 < len(out):
            m2 = re.match(r"\s+/\* (0x[0-9a-f]+) \*/", out[i + 1])
            if m2:
                hi = int(m2.group(1), 16); i += 1
        funcs[cur].append((int(m.group(1), 16), m.group(2), hi))
    i += 1
for name, ins in funcs.items():
    if sys.argv[2] not in name:
        continue
    addr = {a: k for k, (a, _, _) in enumerate(ins)}
    print(name, len(ins), "instructions")
    for k, (a, t, hi) in enumerate(ins):
        m = re.search(r"BRA\S*\s+(?:\S+,\s*)?0x([0-9a-f]+)", t)
        if m and int(m.group(1), 16) < a and int(m.group(1), 16) in addr:
            body = ins[addr[int(m.group(1), 16)]:k + 1]
            if len(body) >= 100:
                print(f"  loop {int(m.group(1), 16):#x}..{a:#x}: {len(body)} instructions, stall sum {sum((x[2] >> 41) & 0xf for x in body)}, "
                      f"FFMA2 {sum('FFMA2' in x[1] for x in body)}, local ld/st {sum(('LDL' in x[1] or 'STL' in x[1]) for x in body)}")
