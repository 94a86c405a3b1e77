#!/bin/bash
# First GPU call of the next round (run from the repo root under gpurun; everything lands in gpurun_out/):
#   1. the whole GPU suite, reporting the tests that were written after round 1's GPU budget was spent (xfail -> XPASS expected);
#   2. cost of the reference bundle order for `merged` (decides whether it becomes the default);
#   3. the headline bench for a before/after reference.
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -rxX 2>&1 | tail -40 > gpurun_out/gpu_suite_rxX.log
for order in canonical libstdcxx; do
  for wl in merged5 merged2; do
    timeout 300 python bench.py --workload $wl --steps 40 --warmup 5 --no-cpu-baseline --merged-bundle-order $order \
      > gpurun_out/bench_${wl}_${order}.json 2> gpurun_out/bench_${wl}_${order}.err
  done
done
for mode in 1 2; do
  for wl in merged5 merged2; do
    timeout 300 python bench.py --workload $wl --steps 40 --warmup 5 --no-cpu-baseline --hot-voxels $mode \
      > gpurun_out/bench_${wl}_hot${mode}.json 2> gpurun_out/bench_${wl}_hot${mode}.err
  done
done
timeout 300 python bench.py --steps 100 --warmup 10 --no-cpu-baseline > gpurun_out/bench_fast5.json 2> gpurun_out/bench_fast5.err
tail -5 gpurun_out/gpu_suite_rxX.log
grep -h -o '"value": [0-9.]*' gpurun_out/bench_*.json | head -8
