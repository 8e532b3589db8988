#!/bin/bash
timeout 600 python -m pytest tests/test_gpu_nodechain.py -q -x > gpurun_out/r02n_pytest.log 2>&1; echo "pytest nodechain rc=$?"; tail -3 gpurun_out/r02n_pytest.log
timeout 900 python -m pytest tests/test_gpu_schnet.py tests/test_gpu_config_parity.py tests/test_gpu_tc.py -q -x -k "schnet or config2 or graphed or cfconv or interaction" -rP > gpurun_out/r02n_pytest2.log 2>&1; echo "pytest schnet rc=$?"; grep -E "passed|failed|^FAILED|^E  " gpurun_out/r02n_pytest2.log | tail -8; grep -h "^.n\[" gpurun_out/r02n_pytest2.log | tail -4
python scripts/prof_nodechain.py 2>&1 | tee gpurun_out/r02n_chain.log
timeout 600 python bench.py --steps 10 --warmup 3 --only 2 > gpurun_out/r02n_bench.json 2> gpurun_out/r02n_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/r02n_bench.err
python - <<'PY'
import json
s = open("gpurun_out/r02n_bench.json").read()
d = json.loads(s[s.find('{"metric'):].splitlines()[0])
print("value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], "roof", d["roofline"]["frac"], "launches", d.get("gpu_launches"), d.get("graph_kernel_nodes"))
PY
