#!/bin/bash
# launch list of the MACE product block (symmetric contraction + o3.Linear + sc) forward + backward at N = 65536
python scripts/bench_layers.py symc > gpurun_out/r02j_plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -s 0 -c 400 --csv --log-file gpurun_out/r02j_launches.csv python scripts/bench_layers.py symc > gpurun_out/r02j_ncu.log 2>&1
cat gpurun_out/r02j_plain.log | tail -1
python scripts/ncu_agg.py gpurun_out/r02j_launches.csv 14
