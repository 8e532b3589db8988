#!/bin/bash
# 2 GPUs: partition tests in every halo mode, then config 5 at N = 2 in the three modes
timeout 900 python -m pytest tests/test_gpu_multi.py -q -x > gpurun_out/r02y_pytest.log 2>&1; echo "pytest rc=$?"; grep -E "passed|failed|^E  |Error" gpurun_out/r02y_pytest.log | tail -8
for H in fused peer; do
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 --only 5 --halo $H > gpurun_out/r02y_bench2_$H.json 2> gpurun_out/r02y_bench2_$H.err; echo "bench $H rc=$?"
tail -c 300 gpurun_out/r02y_bench2_$H.err
python - <<PY
import json
s = open("gpurun_out/r02y_bench2_$H.json").read()
d = json.loads(s[s.find('{"metric'):].splitlines()[0])
for k, v in d["configs"].items():
    print(k, v.get("ms_per_step"), v.get("edges_per_s_per_layer"), str(v.get("halo_exchange"))[:90], v.get("partition_check"), v.get("error"), v.get("trace"))
PY
done
