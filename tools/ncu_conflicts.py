"""Per-source-line shared-memory wavefronts: excess over the ideal count (= bank-conflict replays) from an ncu
report taken with --set full --import-source on.  Usage: python tools/ncu_conflicts.py report.ncu-rep [top_n]"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
hdr, cur, out = None, None, []
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur = r[1].split("/")[-1]
        continue
    if r[0] == "Line No":
        hdr = r
        if "--cols" in sys.argv:
            print(hdr)
        continue
    if hdr is None or not r[0].isdigit():
        continue
    def col(name):
        try:
            return float(r[hdr.index(name)] or 0)
        except (ValueError, IndexError):
            return 0.0
    w, ideal = col("L1 Wavefronts Shared"), col("L1 Wavefronts Shared Ideal")
    if w > 0:
        out.append((w - ideal, w, ideal, cur, int(r[0]), r[1].strip()[:100]))
tot_w, tot_x = sum(o[1] for o in out), sum(o[0] for o in out)
print(f"shared wavefronts {tot_w:,.0f}, excess over ideal {tot_x:,.0f} ({100 * tot_x / max(tot_w, 1):.1f} %)")
for x, w, ideal, f, ln, src in sorted(out, reverse=True)[:top]:
    print(f"{100 * x / max(tot_x, 1):5.1f}% of excess  {x:14,.0f} / {w:14,.0f}  {f}:{ln:<4d} {src}")
