"""Per-source-line instruction counts / stall samples from an ncu report (needs -lineinfo).
Usage: python tools/ncu_lines.py report.ncu-rep [top_n]"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
cur_file, hdr, lines = None, None, []
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
        continue
    if r[0] == "Line No":
        hdr = r
        continue
    if hdr is None or r[0] in ("Function Name",):
        continue
    if r[0].isdigit():
        try:
            ie = int(r[hdr.index("Instructions Executed")])
            ss = int(r[hdr.index("# Samples")])
        except (ValueError, IndexError):
            continue
        lines.append((ie, ss, cur_file, int(r[0]), r[1].strip()[:110]))
tot_i = sum(l[0] for l in lines) or 1
tot_s = sum(l[1] for l in lines) or 1
print(f"total warp-instructions {tot_i:,}  samples {tot_s:,}")
print("--- by instructions executed")
for ie, ss, f, ln, src in sorted(lines, reverse=True)[:top]:
    print(f"{100*ie/tot_i:5.1f}% inst {100*ss/tot_s:5.1f}% smpl  {f}:{ln:<4d} {src}")
print("--- by stall samples")
for ie, ss, f, ln, src in sorted(lines, key=lambda l: -l[1])[:top // 2]:
    print(f"{100*ie/tot_i:5.1f}% inst {100*ss/tot_s:5.1f}% smpl  {f}:{ln:<4d} {src}")
