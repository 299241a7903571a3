"""Stall-sample share per source-line range (ncu report, -lineinfo).
Usage: python tools/ncu_phase.py report.ncu-rep file.cu name:lo-hi ..."""
import csv, io, subprocess, sys
rep, fname = sys.argv[1], sys.argv[2]
ranges = []
for spec in sys.argv[3:]:
    n, r = spec.split(":"); lo, hi = r.split("-"); ranges.append((n, int(lo), int(hi)))
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
hdr = None; curfile = None; seen = set()
tot_s = tot_i = 0; acc = {n: [0, 0] for n, _, _ in ranges}; other = [0, 0]
for r in rows:
    if not r: continue
    if r[0] == "File Path": curfile = r[1].split("/")[-1]; continue
    if r[0] == "Line No": hdr = r; continue
    if hdr is None or not r[0].isdigit(): continue
    key = (curfile, int(r[0]))
    if key in seen: continue
    seen.add(key)
    try:
        ss = int(r[hdr.index("# Samples")]); ie = int(r[hdr.index("Instructions Executed")])
    except ValueError:
        continue
    tot_s += ss; tot_i += ie
    hit = False
    if curfile == fname:
        for n, lo, hi in ranges:
            if lo <= key[1] <= hi: acc[n][0] += ss; acc[n][1] += ie; hit = True; break
    if not hit: other[0] += ss; other[1] += ie
for n, (ss, ie) in acc.items(): print(f"{n:14s} {100*ss/tot_s:5.1f}% samples {100*ie/tot_i:5.1f}% inst")
print(f"{'other':14s} {100*other[0]/tot_s:5.1f}% samples {100*other[1]/tot_i:5.1f}% inst   (total samples {tot_s:,}, inst {tot_i:,})")
