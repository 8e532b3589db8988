#!/bin/bash
# full GPU suite + the driver's bench command at N = 1
SECONDS=0; timeout 1500 python bench.py > gpurun_out/r02s_bench.json 2> gpurun_out/r02s_bench.err; echo "bench rc=$? wall ${SECONDS}s"
python - <<'PY'
import json
s = open("gpurun_out/r02s_bench.json").read()
d = json.loads(s[s.find('{"metric'):].splitlines()[0])
print("value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], "roof", d["roofline"]["frac"], {k: (round(v["ms_per_launch"], 4), round(v["frac"], 3)) for k, v in d["roofline"]["both"].items()}, "launches", d.get("gpu_launches"), d.get("graph_kernel_nodes"))
print("cpu", d["cpu_baseline"], "strict", d["fp32_strict"])
for k, v in d["configs"].items():
    print(k, v.get("ms_per_step"), v.get("edges_per_s_per_layer"), v.get("roofline", {}).get("frac"), v.get("hidden_layer", {}).get("ms_fwd_bwd"), v.get("error"))
PY
