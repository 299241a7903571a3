"""Dynamic SASS opcode mix (warp-level executed instructions) and stall samples per opcode from an ncu report.
Usage: python tools/ncu_opmix.py report.ncu-rep [kernel-index]"""
import csv, io, subprocess, sys, collections, re
rep = sys.argv[1]
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
kern = -1; want = int(sys.argv[2]) if len(sys.argv) > 2 else 0
hdr = None; ops = collections.Counter(); smp = collections.Counter()
for r in rows:
    if not r: continue
    if r[0] == "Kernel Name": kern += 1; continue
    if r[0] == "Address": hdr = r; continue
    if kern != want or hdr is None or not r[0].startswith("0x"): continue
    src = r[1].strip()
    m = re.match(r"(@!?U?P\w+\s+)?([A-Z0-9_]+)", src)
    op = m.group(2) if m else "?"
    ops[op] += int(r[hdr.index("Instructions Executed")]); smp[op] += int(r[hdr.index("# Samples")])
ti, ts = sum(ops.values()) or 1, sum(smp.values()) or 1
print(f"total warp-instr {ti:,} samples {ts:,}")
for op, n in ops.most_common(30):
    print(f"{op:10s} {100*n/ti:5.1f}% inst  {100*smp[op]/ts:5.1f}% samples")
