#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_egnn.py tests/test_gpu_tfn.py -q > gpurun_out/r03a_pytest.log 2>&1; echo "pytest rc=$?"; grep -E "passed|failed|^E  |^FAILED" gpurun_out/r03a_pytest.log | tail -8
