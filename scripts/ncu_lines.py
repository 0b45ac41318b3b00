#!/usr/bin/env python
"""Per-source-line summary of an ncu report: python scripts/ncu_lines.py rep.ncu-rep [min_pct]
(runs `ncu -i rep --page source --print-source cuda,sass --csv` and keeps the CUDA-line rows)."""
import csv, subprocess, sys, io
rep = sys.argv[1]; minp = float(sys.argv[2]) if len(sys.argv) > 2 else 0.7
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
hdr = None; fn = None; out = []
for r in rows:
    if r and r[0] == "File Path": fn = r[1].split("/")[-1]; continue
    if r and r[0] == "Line No": hdr = r; continue
    if hdr and len(r) == len(hdr) and r[0].isdigit():
        d = {}
        for k, v in zip(hdr, r):
            if k not in d: d[k] = v
        out.append((fn, int(r[0]), r[1], d))
tot = sum(int(o[3]["Instructions Executed"] or 0) for o in out)
tots = sum(int(o[3]["# Samples"] or 0) for o in out)
print(f"total warp instructions {tot}, samples {tots}")
for fn, ln, src, d in out:
    n = int(d["Instructions Executed"] or 0); s = int(d["# Samples"] or 0)
    if n > tot * minp / 100 or s > tots * minp / 100:
        print(f"{fn}:{ln:4d} inst {100*n/tot:5.1f}% samp {100*s/tots:5.1f}% thr/inst {d['Avg. Threads Executed']:>3} | {src.strip()[:100]}")
if len(sys.argv) > 3:   # region sums: "name:lo-hi,name:lo-hi" on the main file
    main = sys.argv[4] if len(sys.argv) > 4 else out[0][0]
    print("regions:")
    acc = {}
    for spec in sys.argv[3].split(","):
        name, rng = spec.split(":"); lo, hi = map(int, rng.split("-"))
        n = sum(int(o[3]["Instructions Executed"] or 0) for o in out if o[0] == main and lo <= o[1] <= hi)
        s = sum(int(o[3]["# Samples"] or 0) for o in out if o[0] == main and lo <= o[1] <= hi)
        print(f"  {name:12s} {lo:4d}-{hi:4d} inst {100*n/tot:5.1f}% samp {100*s/tots:5.1f}%")
    n = sum(int(o[3]["Instructions Executed"] or 0) for o in out if o[0] != main)
    s = sum(int(o[3]["# Samples"] or 0) for o in out if o[0] != main)
    print(f"  other files inst {100*n/tot:5.1f}% samp {100*s/tots:5.1f}%")
