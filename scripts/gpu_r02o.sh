#!/bin/bash
timeout 600 python -m pytest tests/test_gpu_nodechain.py -q -x > gpurun_out/r02o_pytest.log 2>&1; echo "pytest nodechain rc=$?"; tail -3 gpurun_out/r02o_pytest.log
python scripts/prof_nodechain.py 2>&1 | tee gpurun_out/r02o_chain.log
timeout 900 python -m pytest tests/test_gpu_egnn.py tests/test_gpu_tc.py tests/test_gpu_config_parity.py tests/test_gpu_schnet.py -q -rP > gpurun_out/r02o_pytest2.log 2>&1; echo "pytest egnn/tc/config rc=$?"; grep -E "passed|failed|^FAILED|^E  " gpurun_out/r02o_pytest2.log | tail -12; grep -h "^.n\[" gpurun_out/r02o_pytest2.log | tail -12
python scripts/prof_egnn.py 18 relu 3 2>&1 | tail -1 | tee gpurun_out/r02o_egnn.log
GMP_EGNN_NODE_CHAIN=0 python scripts/prof_egnn.py 18 relu 3 2>&1 | tail -1 | tee -a gpurun_out/r02o_egnn.log
