#!/bin/bash
# final single-GPU verification: smoke(), full GPU suite, the driver's bench command, ncu launch list of the same command
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -4
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r03b_pytest.log 2>&1; echo "pytest rc=$?"; grep -E "passed|failed|^FAILED|^E  " gpurun_out/r03b_pytest.log | tail -6
SECONDS=0; timeout 1500 python bench.py > gpurun_out/r03b_bench.json 2> gpurun_out/r03b_bench.err; echo "bench rc=$? wall ${SECONDS}s"
python - <<'PY'
import json
s = open("gpurun_out/r03b_bench.json").read()
d = json.loads(s[s.find('{"metric'):].splitlines()[0])
print("value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], "roof", d["roofline"]["frac"], "traffic", d["roofline"]["traffic"], "launches", d.get("gpu_launches"), d.get("graph_kernel_nodes"), d.get("clocks"))
for k, v in d["configs"].items():
    print(k, v.get("ms_per_step"), v.get("edges_per_s_per_layer"), v.get("roofline", {}).get("frac"), v.get("hidden_layer", {}).get("ms_fwd_bwd"), v.get("error"))
PY
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r03b_launches_bench.csv python bench.py --only 2 --steps 2 --warmup 3 --no-cpu-baseline --no-strict > gpurun_out/r03b_ncu.log 2>&1; echo "ncu rc=$?"
python scripts/ncu_agg.py gpurun_out/r03b_launches_bench.csv 14
