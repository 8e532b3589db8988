#!/bin/bash
timeout 600 python -m pytest tests/test_gpu_gvp.py tests/test_gpu_mace_blocks.py -q -k equivariance > gpurun_out/r03e_pytest.log 2>&1; echo "pytest rc=$?"; grep -E "passed|failed|^E  |^FAILED" gpurun_out/r03e_pytest.log | tail -8
