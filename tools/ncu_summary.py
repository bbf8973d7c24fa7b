"""Summarise an .ncu-rep (read here, on the CPU box) into the JSON kept under profiles/.
usage: python tools/ncu_summary.py report.ncu-rep "command line that produced it" "note" > profiles/<name>.json
Keeps one record per launch: duration, DRAM bytes, issue / tensor / occupancy figures."""
import csv
import json
import subprocess
import sys

rep, command, note = sys.argv[1], sys.argv[2], sys.argv[3]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, body = rows[0], rows[1], rows[2:]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__cluster_size" if "launch__cluster_size" in hdr else "launch__grid_size",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed"]
out = {"command": command, "note": note, "launches": []}
for r in body:
    rec = {"kernel": r[hdr.index("Kernel Name")].replace("void bfb::<", "").replace("void ", "")[:90]}
    for w in want:
        if w in hdr:
            i = hdr.index(w)
            rec[w] = f"{r[i]} {units[i]}".strip()
    out["launches"].append(rec)


def to_bytes(v):
    num, unit = v.split()[0], (v.split() + [""])[1].lower()
    return float(num) * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(unit, 1)


tot = [to_bytes(l["dram__bytes_read.sum"]) + to_bytes(l["dram__bytes_write.sum"]) for l in out["launches"]
       if "dram__bytes_read.sum" in l]
if tot:
    out["dram_bytes_per_launch"] = tot
    out["dram_bytes_per_launch_mean"] = sum(tot) / len(tot)
print(json.dumps(out, indent=1))
