#!/bin/bash
# 2 GPUs: the NCCL partition test, then the bench line at N = 2
timeout 300 python -m pytest tests/test_gpu_multi.py -q -x > gpurun_out/r02h_pytest.log 2>&1; echo "pytest multi rc=$?"; tail -3 gpurun_out/r02h_pytest.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r02h_bench2.json 2> gpurun_out/r02h_bench2.err; echo "bench rc=$?"
tail -c 600 gpurun_out/r02h_bench2.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/r02h_bench2.json"))
print("value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"])
for k, v in d["configs"].items():
    print(k, v.get("ms_per_step"), v.get("edges_per_s_per_layer"), v.get("roofline", {}).get("frac"), v.get("partition_check"), v.get("halo_nodes_max"), v.get("error"), v.get("trace"))
PY
