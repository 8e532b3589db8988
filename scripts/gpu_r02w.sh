#!/bin/bash
# mbarrier wait-loop variants: rebuild with a fixed sleep between probes and time the pipelined kernels
for NS in 0 20 50 100 200; do
  if [ $NS -eq 0 ]; then unset GMP_NVCC_EXTRA; else export GMP_NVCC_EXTRA=-DGMP_MBAR_SLEEP_NS=$NS; fi
  python geometric-message-passing_b200/build.py --force > /dev/null 2>&1 || echo BUILD FAILED
  echo "== sleep $NS ns"
  for k in schnet_fwd2k schnet_bwd2; do timeout 120 python scripts/prof_kernel.py $k bf16 10; done 2>&1 | grep ms/launch
  python scripts/prof_egnn.py 18 relu 3 2>&1 | tail -1
  python scripts/bench_layers.py tfn 2>&1 | tail -1 | cut -c1-250
done 2>&1 | tee gpurun_out/r02w_sleep.log
unset GMP_NVCC_EXTRA
