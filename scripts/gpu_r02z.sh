#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_multi.py -q > gpurun_out/r02z_pytest.log 2>&1; echo "pytest rc=$?"; grep -E "passed|failed|^E  |Error" gpurun_out/r02z_pytest.log | tail -8
