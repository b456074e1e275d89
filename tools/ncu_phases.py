#!/usr/bin/env python3
"""Per-phase breakdown of an `ncu --set full --import-source on` capture of a step kernel (dev tool).
usage: tools/ncu_phases.py report.ncu-rep [n_top_lines]
Exports the cuda,sass source page, orders the SASS by address and reports warp-stall samples / executed instructions
per source line and per address bucket (the step's phases are laid out in program order)."""
import collections
import csv
import subprocess
import sys


def load(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    cur = hdr = curline = None
    recs, src = [], {}
    for r in rows:
        if len(r) == 2 and r[0] == "File Path":
            cur = r[1].split('/')[-1]; continue
        if r and r[0] == "Line No":
            hdr = r; continue
        if hdr is None or len(r) < 10:
            continue
        if r[0] != "":
            curline = (cur, int(r[0])); src[curline] = r[1][:100]; continue
        if r[2].startswith('0x'):
            try:
                recs.append((int(r[2], 16), r[3].strip(), int(r[4]), int(r[7]), curline, r))
            except ValueError:
                pass
    recs.sort(key=lambda x: x[0])
    return hdr, recs, src


def main():
    rep = sys.argv[1]
    ntop = int(sys.argv[2]) if len(sys.argv) > 2 else 30
    hdr, recs, src = load(rep)
    tot = sum(r[2] for r in recs)
    print(f"{len(recs)} SASS instructions, {tot} samples, {sum(r[3] for r in recs)} warp instructions executed")
    si = [i for i, n in enumerate(hdr) if n.startswith('stall_') and 'Not Issued' not in n]
    byline, instl = collections.Counter(), collections.Counter()
    for r in recs:
        byline[r[4]] += r[2]; instl[r[4]] += r[3]
    print("--- top source lines by samples")
    for k, v in byline.most_common(ntop):
        print(f"{k[0]}:{k[1]:<5d} {v:5d} {100 * v / max(tot, 1):5.1f}%  inst {instl[k]:8d}  {src.get(k, '')}")
    print("--- address buckets (400 SASS instructions each)")
    base = recs[0][0]
    for i in range(0, len(recs), 400):
        ch = recs[i:i + 400]
        s = sum(r[2] for r in ch)
        st = collections.Counter()
        for r in ch:
            for j in si:
                try:
                    st[hdr[j][6:]] += int(r[5][j])
                except ValueError:
                    pass
        lines = collections.Counter(r[4] for r in ch).most_common(2)
        print(f"{i:6d} {hex(ch[0][0] - base):>8s} samples {s:5d} {100 * s / max(tot, 1):5.1f}% inst {sum(r[3] for r in ch):8d} "
              f"{[(k, v) for k, v in st.most_common(4)]} {[(f'{k[0]}:{k[1]}') for k, _ in lines]}")


if __name__ == "__main__":
    main()
