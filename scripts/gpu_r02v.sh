#!/bin/bash
# ncu --set full of the two dominant SchNet kernels of the CURRENT build (for bench.py's roofline.traffic) + the segsum + wgrad kernels
python scripts/prof_kernel.py schnet_fwd2k bf16 3 > /dev/null 2>&1 && ncu --set full --clock-control none --import-source on -k regex:schnet_fwd_tc2 -s 2 -c 1 -o gpurun_out/r02v_fwd2k -f python scripts/prof_kernel.py schnet_fwd2k bf16 3 > gpurun_out/r02v_ncu1.log 2>&1
python scripts/prof_kernel.py schnet_bwd2 bf16 3 > /dev/null 2>&1 && ncu --set full --clock-control none --import-source on -k regex:schnet_bwd_tc2 -s 2 -c 1 -o gpurun_out/r02v_bwd2 -f python scripts/prof_kernel.py schnet_bwd2 bf16 3 > gpurun_out/r02v_ncu2.log 2>&1
ls -la gpurun_out/r02v_*.ncu-rep
for k in schnet_fwd2 schnet_fwd2k schnet_bwd2; do timeout 120 python scripts/prof_kernel.py $k bf16 10; done 2>&1 | tee gpurun_out/r02v_kernels.log
