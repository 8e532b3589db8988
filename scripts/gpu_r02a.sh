#!/bin/bash
# GPU round-trip: full GPU test suite, then the default bench line
python -m pytest tests -m gpu -q -x --durations=15 > gpurun_out/r02a_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02a_pytest.log
tail -5 gpurun_out/r02a_pytest.log
free -g | head -2; nproc
python bench.py --steps 10 --warmup 3 > gpurun_out/r02a_bench.json 2> gpurun_out/r02a_bench.err; echo "bench rc=$?"
tail -c 1500 gpurun_out/r02a_bench.err
