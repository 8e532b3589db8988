#!/bin/bash
# the driver's scaling command at N = $1 (all configs), one JSON line
N=${1:-8}
SECONDS=0
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r02u_bench$N.json 2> gpurun_out/r02u_bench$N.err; echo "bench N=$N rc=$? wall ${SECONDS}s"
tail -c 300 gpurun_out/r02u_bench$N.err
python - <<PY
import json
s = open("gpurun_out/r02u_bench$N.json").read()
d = json.loads(s[s.find('{"metric'):].splitlines()[0])
print("value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"])
for k, v in d["configs"].items():
    print(k, v.get("ms_per_step"), v.get("edges_per_s_per_layer"), v.get("roofline", {}).get("frac"), v.get("halo_exchange"), v.get("halo_nodes_max"), v.get("partition_check"), v.get("error"), v.get("trace"))
PY
