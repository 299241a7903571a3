"""profiles/traffic.json from the committed ncu summaries of a round (tools/summarize_ncu.py output):
per bench env, DRAM bytes of one launch of the bench configuration, fp64-pipe and issue-slot activity.
Usage: python tools/make_traffic.py <round-tag>     (reads profiles/<tag>_ncu_<env>.json)"""
import json
import os
import sys

tag = sys.argv[1]
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
out = {}
for env in ("shkadov", "shkadov_separable", "rayleigh", "mixing"):
    p = os.path.join(ROOT, "profiles", f"{tag}_ncu_{env}.json")
    if not os.path.exists(p):
        continue
    l = json.load(open(p))["launches"][0]
    f = lambda k: float(l[k]["value"])
    unit = {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1.0}
    b = sum(f(k) * unit[l[k]["unit"]] for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
    out[env] = {
        "bytes_per_launch": b, "grid": int(f("launch__grid_size")),
        "source": f"profiles/{tag}_ncu_{env}.json (ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum, one launch of the bench configuration)",
        "fp64_pipe_active_pct": f("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"),
        "issue_active_pct": f("smsp__issue_active.avg.pct_of_peak_sustained_active"),
    }
json.dump(out, open(os.path.join(ROOT, "profiles", "traffic.json"), "w"), indent=1)
print(json.dumps(out, indent=1))
