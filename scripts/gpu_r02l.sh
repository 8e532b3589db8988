#!/bin/bash
# uvu interaction blocks: parity tests
timeout 600 python -m pytest tests/test_gpu_mace_blocks.py -q -x > gpurun_out/r02l_pytest.log 2>&1; echo "pytest rc=$?"; tail -25 gpurun_out/r02l_pytest.log
