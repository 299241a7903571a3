"""Summarise an ncu report (.ncu-rep) into a small JSON that can be committed under profiles/.
Usage: python tools/summarize_ncu.py gpurun_out/prof_x.ncu-rep profiles/ncu_x.json"""
import csv
import io
import json
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
    "launch__block_size", "launch__waves_per_multiprocessor", "launch__shared_mem_per_block_dynamic",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "sm__cycles_elapsed.max",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "smsp__sass_inst_executed_op_local_ld.sum", "smsp__sass_inst_executed_op_local_st.sum",
    "smsp__inst_executed_pipe_fp64.sum", "sm__inst_executed_pipe_fp64.sum", "smsp__inst_executed_op_shared_ld.sum",
    "smsp__inst_executed_op_shared_st.sum", "sm__cycles_active.avg", "smsp__cycles_active.avg",
]

rep, out = sys.argv[1], sys.argv[2]
txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
hdr, units = rows[0], rows[1]
res = []
for r in rows[2:]:
    d = {"kernel": r[hdr.index("Kernel Name")]}
    for k in KEYS:
        if k in hdr:
            i = hdr.index(k)
            d[k] = {"value": r[i], "unit": units[i]}
    stalls = {}
    for i, k in enumerate(hdr):
        if "issue_stalled" in k and k.endswith("per_issue_active.ratio"):
            try:
                v = float(r[i])
            except ValueError:
                continue
            if v >= 0.05:
                stalls[k.split("issue_stalled_")[1].replace("_per_issue_active.ratio", "")] = round(v, 3)
    d["stalls_per_issue_active"] = dict(sorted(stalls.items(), key=lambda kv: -kv[1]))
    res.append(d)
json.dump({"report": rep, "command": "ncu --set full --clock-control none --import-source on", "launches": res}, open(out, "w"), indent=1)
print("wrote", out, len(res), "launches")
