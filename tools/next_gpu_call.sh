#!/bin/bash
# GPU call (run from the repo root under gpurun; everything lands in gpurun_out/r02/)
set -u
O=gpurun_out/r02
mkdir -p $O
timeout 120 tools/ubench_launch > $O/ubench_launch.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_ref_golden.py tests/test_gpu_bundle_order.py tests/test_gpu_parity.py -q -m gpu -x 2>&1 | tail -15 > $O/gpu_merged_order.log
run() { name=$1; shift; timeout 400 python bench.py --steps 30 --warmup 5 --no-cpu-baseline "$@" > $O/bench_${name}.json 2> $O/bench_${name}.err; }
run merged2_reforder_v2 --workload merged2
run merged5_reforder_v2 --workload merged5
cat $O/ubench_launch.txt
tail -8 $O/gpu_merged_order.log
for f in $O/bench_*v2.json; do echo "$f: $(grep -o '"value": [0-9.]*' $f | head -1) $(grep -o '"phase_ms_per_frame": {[^}]*}' $f)"; done
