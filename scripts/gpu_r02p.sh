#!/bin/bash
# launch list of the SchNet step (eager, 2 steps: the second is warm) + EGNN config parity re-check
python scripts/prof_step.py 2 > /dev/null 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/r02p_launches.csv python scripts/prof_step.py 2 > gpurun_out/r02p_ncu.log 2>&1
python scripts/ncu_agg.py gpurun_out/r02p_launches.csv 25
timeout 600 python -m pytest tests/test_gpu_config_parity.py -q -k "config5 or config1" -rP > gpurun_out/r02p_pytest.log 2>&1; echo "pytest rc=$?"; grep -E "passed|failed|^FAILED|^E  " gpurun_out/r02p_pytest.log | tail -5
