#!/bin/bash
# GPU call 23: shim (mesh accessor through the drop-in classes) + the last library build
set -u
O=gpurun_out/r02
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_more.py tests/test_gpu_mesh.py tests/test_gpu_parity.py -q -m gpu -x 2>&1 | tail -6 > $O/gpu_quick_23.log
tail -4 $O/gpu_quick_23.log
