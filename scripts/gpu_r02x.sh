#!/bin/bash
timeout 600 python -m pytest tests/test_gpu_nodechain.py tests/test_gpu_schnet.py -q -x > gpurun_out/r02x_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r02x_pytest.log
python scripts/prof_nodechain.py 2>&1 | grep wgrad | tee gpurun_out/r02x_wgrad.log
GMP_WGRAD_V1=1 python scripts/prof_nodechain.py 2>&1 | grep wgrad | tee -a gpurun_out/r02x_wgrad.log
timeout 600 python bench.py --steps 10 --warmup 3 --only 2 --no-cpu-baseline --no-strict > gpurun_out/r02x_bench.json 2> gpurun_out/r02x_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
s = open("gpurun_out/r02x_bench.json").read()
d = json.loads(s[s.find('{"metric'):].splitlines()[0])
print("value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], "roof", d["roofline"]["frac"], "traffic", d["roofline"]["traffic"])
PY
