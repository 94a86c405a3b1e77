#!/bin/bash
# GPU call (run from the repo root under gpurun; everything lands in gpurun_out/r02/)
set -u
O=gpurun_out/r02
mkdir -p $O
timeout 120 tools/ubench_launch > $O/ubench_launch.txt 2>&1
cat $O/ubench_launch.txt
