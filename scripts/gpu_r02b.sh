#!/bin/bash
# GPU round-trip: full GPU test suite (all failures), then ncu --set full of the two dominant SchNet kernels
python -m pytest tests -m gpu -q -rP --durations=10 > gpurun_out/r02b_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02b_pytest.log
grep -E "^\[|passed|failed|FAILED|rc=" gpurun_out/r02b_pytest.log | tail -40
python scripts/prof_kernel.py schnet_fwd2k bf16 3 > gpurun_out/r02b_plain_fwd.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:schnet_fwd_tc2 -s 2 -c 1 -o gpurun_out/r02b_fwd2k -f python scripts/prof_kernel.py schnet_fwd2k bf16 3 > gpurun_out/r02b_ncu_fwd.log 2>&1
python scripts/prof_kernel.py schnet_bwd2 bf16 3 > gpurun_out/r02b_plain_bwd.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:schnet_bwd_tc2 -s 2 -c 1 -o gpurun_out/r02b_bwd2 -f python scripts/prof_kernel.py schnet_bwd2 bf16 3 > gpurun_out/r02b_ncu_bwd.log 2>&1
cat gpurun_out/r02b_plain_fwd.log gpurun_out/r02b_plain_bwd.log | tail -4
tail -3 gpurun_out/r02b_ncu_fwd.log gpurun_out/r02b_ncu_bwd.log
ls -la gpurun_out/*.ncu-rep
