#!/usr/bin/env python
"""Per-source-line executed-instruction histogram from an ncu report captured with --import-source on.
usage: ncu_lines.py report.ncu-rep [top_n] [samples]   (third argument "samples": order by stall samples)"""
import collections
import csv
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
cur = None
hdr = None
agg = collections.Counter()
smp = collections.Counter()
src = {}
last = None
for r in csv.reader(out.splitlines()):
    if len(r) == 2 and r[0] == "File Path":
        cur = r[1].split("/")[-1]
        continue
    if len(r) == 2 and r[0] == "Function Name":
        print("==", r[1])
        continue
    if r and r[0] == "Line No":
        hdr = r
        continue
    if hdr is None or len(r) < 8:
        continue
    if r[0] != "":
        last = (cur, int(r[0]))
        src[last] = r[1]
    if r[2].startswith("0x"):
        try:
            agg[last] += int(r[7])
            smp[last] += int(r[6])
        except ValueError:
            pass
tot = sum(agg.values())
stot = sum(smp.values()) or 1
print("warp instructions executed:", tot)
by = smp if (len(sys.argv) > 3 and sys.argv[3] == "samples") else agg
for k, _ in by.most_common(top):
    print(f"{agg[k] / tot * 100:5.1f}% inst {smp[k] / stot * 100:5.1f}% samples  {k[0]}:{k[1]}  {src[k].strip()[:110]}")
