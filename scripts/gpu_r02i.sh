#!/bin/bash
# ncu --set full of the fused EGNN backward kernel (csrc/egnn_tc.cu) at 2^18 nodes
python scripts/prof_egnn.py 18 relu 1 > gpurun_out/r02i_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:egnn_bwd -s 1 -c 1 -o gpurun_out/r02i_egnn_bwd -f python scripts/prof_egnn.py 18 relu 1 > gpurun_out/r02i_ncu.log 2>&1
tail -1 gpurun_out/r02i_plain.log; tail -3 gpurun_out/r02i_ncu.log; ls -la gpurun_out/r02i_egnn_bwd.ncu-rep
