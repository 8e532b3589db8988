#!/bin/bash
# ncu --set full of the round's new kernels in their final form: node chain (SchNet forward chain), pipelined wgrad, uvu forward
ncu --set full --clock-control none --import-source on -k regex:"node_chain_kernel|linear_wgrad_tc3" -s 6 -c 3 -o gpurun_out/r03l_node -f python scripts/prof_step.py 1 > gpurun_out/r03l_ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:uvu_conv_fwd -s 1 -c 1 -o gpurun_out/r03l_uvu -f python scripts/bench_layers.py uvu > gpurun_out/r03l_ncu2.log 2>&1
ls -la gpurun_out/r03l_*.ncu-rep
