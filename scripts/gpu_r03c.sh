#!/bin/bash
timeout 600 python -m pytest tests/test_gpu_gvp.py -q > gpurun_out/r03c_pytest.log 2>&1; echo "pytest rc=$?"; grep -E "passed|failed|^E  |^FAILED" gpurun_out/r03c_pytest.log | tail -8
