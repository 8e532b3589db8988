#!/bin/bash
# SchNet kernel changes: carry-based read-out (fwd), bf16 g rows + 8-warp epiP (bwd)
timeout 600 python -m pytest tests/test_gpu_tc.py tests/test_gpu_schnet.py tests/test_gpu_config_parity.py -q -x -k "cfconv or schnet or graphed or interaction or config2 or config3 or config4 or gather" -rP > gpurun_out/r02d_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02d_pytest.log
grep -E "passed|failed|^FAILED|rc=|^E  " gpurun_out/r02d_pytest.log | tail -12
grep -h "^.n\[" gpurun_out/r02d_pytest.log | tail -8
for k in schnet_fwd2 schnet_fwd2k schnet_bwd2; do timeout 120 python scripts/prof_kernel.py $k bf16 10; done 2>&1 | tee gpurun_out/r02d_kernels.log
timeout 300 python bench.py --steps 10 --warmup 3 --only 2 --no-cpu-baseline > gpurun_out/r02d_bench.json 2> gpurun_out/r02d_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d = json.load(open("gpurun_out/r02d_bench.json"))
print("value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], "roof", d["roofline"]["frac"], {k: (v["ms_per_launch"], v["frac"]) for k, v in d["roofline"]["both"].items()})
PY
# EGNN: new thread-per-row forward (csrc/egnn_tc2.cu)
timeout 900 python -m pytest tests/test_gpu_tc.py tests/test_gpu_egnn.py tests/test_gpu_config_parity.py -q -k "egnn" -rP > gpurun_out/r02d_pytest_egnn.log 2>&1; echo "pytest egnn rc=$?" >> gpurun_out/r02d_pytest_egnn.log
grep -E "passed|failed|^FAILED|rc=|^E  " gpurun_out/r02d_pytest_egnn.log | tail -12
grep -h "^.n\[" gpurun_out/r02d_pytest_egnn.log | tail -8
timeout 200 python scripts/prof_egnn.py 18 relu 5 2>&1 | tail -2
GMP_EGNN_TC2=0 timeout 200 python scripts/prof_egnn.py 18 relu 5 2>&1 | tail -2
