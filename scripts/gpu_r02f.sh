#!/bin/bash
# SchNet bwd variant B2 timing + ncu --set full of: SchNet fwd (keep), SchNet bwd, EGNN fwd tc2, EGNN bwd (fused)
timeout 300 python -m pytest tests/test_gpu_tc.py tests/test_gpu_schnet.py -q -x -k "cfconv or schnet or graphed or interaction" > gpurun_out/r02f_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r02f_pytest.log
for k in schnet_fwd2k schnet_bwd2; do timeout 120 python scripts/prof_kernel.py $k bf16 10; done 2>&1 | tee gpurun_out/r02f_kernels.log
python scripts/prof_kernel.py schnet_fwd2k bf16 3 > /dev/null 2>&1 && ncu --set full --clock-control none --import-source on -k regex:schnet_fwd_tc2 -s 2 -c 1 -o gpurun_out/r02f_fwd2k -f python scripts/prof_kernel.py schnet_fwd2k bf16 3 > gpurun_out/r02f_ncu1.log 2>&1
python scripts/prof_kernel.py schnet_bwd2 bf16 3 > /dev/null 2>&1 && ncu --set full --clock-control none --import-source on -k regex:schnet_bwd_tc2 -s 2 -c 1 -o gpurun_out/r02f_bwd2 -f python scripts/prof_kernel.py schnet_bwd2 bf16 3 > gpurun_out/r02f_ncu2.log 2>&1
python scripts/prof_egnn.py 18 relu 1 > gpurun_out/r02f_plain_egnn.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:egnn_ -s 6 -c 2 -o gpurun_out/r02f_egnn -f python scripts/prof_egnn.py 18 relu 1 > gpurun_out/r02f_ncu3.log 2>&1
cat gpurun_out/r02f_plain_egnn.log | tail -1; ls -la gpurun_out/r02f_*.ncu-rep
