#!/bin/bash
timeout 600 python -m pytest tests/test_gpu_nodechain.py -q -x > gpurun_out/r02m_pytest.log 2>&1; echo "pytest rc=$?"; tail -25 gpurun_out/r02m_pytest.log
