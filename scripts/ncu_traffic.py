"""profiles/r02_ncu_traffic.json from `ncu --set full` reports: per kernel, DRAM bytes per launch (read + write), duration and
the counters bench.py's roofline note cites.   usage: python scripts/ncu_traffic.py <report.ncu-rep> [...]"""
import csv
import io
import json
import os
import subprocess
import sys

WANT = {"dram__bytes_read.sum": "dram_read", "dram__bytes_write.sum": "dram_write", "gpu__time_duration.sum": "duration",
        "l1tex__data_pipe_lsu_wavefronts.sum.pct_of_peak_sustained_elapsed": "lsu_wavefronts_pct",
        "sm__inst_executed.avg.per_cycle_active": "ipc", "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm_throughput_pct",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_throughput_pct", "launch__registers_per_thread": "registers"}
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.0}
out = {}
for rep in sys.argv[1:]:
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        u = dict(zip(hdr, units))
        name = d["Kernel Name"].split("(")[0].split("::")[-1].replace("void ", "").strip()
        rec = {"report": os.path.basename(rep)}
        for k, short in WANT.items():
            if k in d and d[k] != "":
                v = float(d[k].replace(",", ""))
                rec[short] = v * UNIT.get(u[k], 1.0) if u[k] in UNIT else v
        if "dram_read" in rec and "dram_write" in rec:
            rec["dram_bytes"] = rec["dram_read"] + rec["dram_write"]
        out.setdefault(name, rec)
json.dump(out, open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "r02_ncu_traffic.json"), "w"), indent=1)
print(json.dumps(out, indent=1))
