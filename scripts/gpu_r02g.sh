#!/bin/bash
# forward: meta block published before the row-stage wait; full GPU suite; bench
for k in schnet_fwd2 schnet_fwd2k schnet_bwd2; do timeout 120 python scripts/prof_kernel.py $k bf16 10; done 2>&1 | tee gpurun_out/r02g_kernels.log
timeout 1200 python -m pytest tests -m gpu -q -rP > gpurun_out/r02g_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02g_pytest.log
grep -E "passed|failed|^FAILED|rc=|^E  " gpurun_out/r02g_pytest.log | tail -12
grep -h "^.n\[" gpurun_out/r02g_pytest.log | tail -12
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r02g_bench.json 2> gpurun_out/r02g_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d = json.load(open("gpurun_out/r02g_bench.json"))
print("value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], "roof", d["roofline"]["frac"], {k: (v["ms_per_launch"], v["frac"]) for k, v in d["roofline"]["both"].items()})
for k, v in d["configs"].items():
    print(k, v.get("ms_per_step"), v.get("edges_per_s_per_layer"), v.get("roofline", {}).get("frac"), v.get("hidden_layer", {}).get("ms_fwd_bwd"), v.get("error"))
PY
