#!/bin/bash
# ncu --set full of the EGNN tcgen05 edge kernels at the config-5 geometry (2^18 nodes)
python scripts/prof_egnn.py 18 relu 3 > gpurun_out/r02e_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:egnn_ -s 6 -c 2 -o gpurun_out/r02e_egnn -f python scripts/prof_egnn.py 18 relu 1 > gpurun_out/r02e_ncu.log 2>&1
cat gpurun_out/r02e_plain.log; tail -n 5 gpurun_out/r02e_ncu.log; ls -la gpurun_out/r02e_egnn.ncu-rep
