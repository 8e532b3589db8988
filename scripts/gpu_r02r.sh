#!/bin/bash
timeout 600 python -m pytest tests/test_gpu_nodechain.py -q -x > gpurun_out/r02r_pytest.log 2>&1; echo "pytest nodechain rc=$?"; tail -2 gpurun_out/r02r_pytest.log
python scripts/prof_nodechain.py 2>&1 | tee gpurun_out/r02r_chain.log
