#!/bin/bash
timeout 600 python -m pytest tests/test_gpu_mace.py tests/test_gpu_config_parity.py -q -k "mace or symmetric or product or tfn" -rP > gpurun_out/r02k_pytest.log 2>&1; echo "pytest rc=$?"
grep -E "passed|failed|^FAILED|^E  " gpurun_out/r02k_pytest.log | tail -8; grep -h "^.n\[" gpurun_out/r02k_pytest.log | tail -6
python scripts/bench_layers.py symc > gpurun_out/r02k_plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -s 0 -c 400 --csv --log-file gpurun_out/r02k_launches.csv python scripts/bench_layers.py symc > gpurun_out/r02k_ncu.log 2>&1
cat gpurun_out/r02k_plain.log | tail -1
python scripts/ncu_agg.py gpurun_out/r02k_launches.csv 8
