"""SASS of a source-line range with per-instruction executed counts / stall samples (ncu report, -lineinfo).
Usage: python tools/ncu_sass_lines.py report.ncu-rep file.cu first last"""
import csv, io, subprocess, sys, re, collections
rep, fname, lo, hi = sys.argv[1], sys.argv[2], int(sys.argv[3]), int(sys.argv[4])
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
# the cuda,sass view lists, per source line, the SASS instructions correlated with it
hdr = None; cur = None; curfile = None
ops = collections.Counter(); tot = 0
for r in rows:
    if not r: continue
    if r[0] == "File Path": curfile = r[1].split("/")[-1]; continue
    if r[0] in ("Line No", "Address"): hdr = r; continue
    if hdr is None: continue
    if r[0].isdigit() and hdr[0] == "Line No":
        cur = int(r[0]); continue
    if r[0] == "" and len(r) > 3 and r[2].startswith("0x") and curfile == fname and cur is not None and lo <= cur <= hi:
        ie = int(r[hdr.index("Instructions Executed")]); ss = int(r[hdr.index("# Samples")])
        m = re.match(r"(@!?U?P\w+\s+)?([A-Z0-9_]+)", r[3].strip())
        ops[m.group(2) if m else "?"] += ie; tot += ie
        if "-v" in sys.argv: print(cur, r[3].strip()[:70], ie, ss)
print("total", tot)
for k, v in ops.most_common(25): print(f"{k:10s} {v:14,d} {100*v/max(tot,1):5.1f}%")
