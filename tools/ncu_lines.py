#!/usr/bin/env python
"""Stall samples per CUDA source line from `ncu -i rep --page source --csv --print-source cuda,sass`.
    python tools/ncu_lines.py rep.ncu-rep [top]"""
import csv, subprocess, sys, io
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
cur = None; data = []
for r in rows:
    if len(r) >= 2 and r[0] == "File Path": cur = r[1].split("/")[-1]; continue
    if len(r) > 8 and r[0].isdigit():
        try: data.append((int(r[6]), int(r[7]), cur, int(r[0]), r[1].strip()))
        except ValueError: pass
tot = sum(d[0] for d in data)
print("total samples", tot)
for d in sorted(data, reverse=True)[:top]:
    print("%6d %5.1f%% inst=%9d  %s:%d: %s" % (d[0], 100.0 * d[0] / tot, d[1], d[2], d[3], d[4][:120]))
