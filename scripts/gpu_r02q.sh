#!/bin/bash
ncu --set full --clock-control none --import-source on -k regex:node_chain_kernel -s 3 -c 2 -o gpurun_out/r02q_chain -f python scripts/prof_step.py 1 > gpurun_out/r02q_ncu.log 2>&1
tail -3 gpurun_out/r02q_ncu.log; ls -la gpurun_out/r02q_chain.ncu-rep
