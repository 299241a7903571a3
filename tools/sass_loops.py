"""Static SASS loop census of one kernel: for every backward branch, the instruction count and
opcode classes of the loop body (largest loops first).  No GPU needed.
Usage: python tools/sass_loops.py file.o <mangled-kernel-substring> [min-instr]"""
import collections, re, subprocess, sys
obj, pat = sys.argv[1], sys.argv[2]
mn = int(sys.argv[3]) if len(sys.argv) > 3 else 200
names = subprocess.run(["cuobjdump", "-elf", obj], capture_output=True, text=True).stdout
txt = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
ins, on = [], False
for line in txt.splitlines():
    if "Function :" in line:
        on = pat in line
        continue
    if not on: continue
    m = re.match(r"\s+/\*([0-9a-f]+)\*/\s+(.*?);", line)
    if m: ins.append((int(m.group(1), 16), m.group(2).strip()))
addr = {a: i for i, (a, _) in enumerate(ins)}
FP64 = ("DFMA", "DMUL", "DADD", "DSETP", "DMNMX")
loops = []
for i, (a, s) in enumerate(ins):
    m = re.search(r"\bBRA(?:\.U)?\s+.*?(0x[0-9a-f]+)", s)
    if m:
        t = int(m.group(1), 16)
        if t <= a and t in addr: loops.append((addr[t], i))
print(f"{len(ins)} instructions, {len(loops)} backward branches")
for lo, hi in sorted(loops, key=lambda l: l[0] - l[1]):
    n = hi - lo + 1
    if n < mn: continue
    ops = collections.Counter()
    for _, s in ins[lo:hi + 1]:
        m = re.match(r"(@!?U?P\w+\s+)?([A-Z0-9_]+)", s)
        ops[m.group(2) if m else "?"] += 1
    f64 = sum(v for k, v in ops.items() if k in FP64)
    print(f"loop [{ins[lo][0]:#x}..{ins[hi][0]:#x}] {n} instr, fp64 {f64}, BAR {ops['BAR']}: " +
          ", ".join(f"{k} {v}" for k, v in ops.most_common(14)))
