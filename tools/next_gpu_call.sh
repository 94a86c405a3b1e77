#!/bin/bash
# GPU call (run from the repo root under gpurun; everything lands in gpurun_out/r02/)
set -u
O=gpurun_out/r02
mkdir -p $O
timeout 180 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke_s3.log 2>&1; echo "smoke rc=$?" >> $O/smoke_s3.log
tail -4 $O/smoke_s3.log
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_ref_golden.py tests/test_gpu_more.py tests/test_gpu_fuzz.py -q -m gpu -x -k "fast or fuzz or reset or shard or import or resume or shim" 2>&1 | tail -15 > $O/gpu_fast_s3.log
tail -6 $O/gpu_fast_s3.log
run() { name=$1; shift; timeout 400 python bench.py --no-cpu-baseline "$@" > $O/bench_${name}.json 2> $O/bench_${name}.err; }
run fast5_s3 --steps 100 --warmup 10
KSG_SOLVER=2 timeout 400 python bench.py --no-cpu-baseline --steps 100 --warmup 10 > $O/bench_fast5_s2.json 2> $O/bench_fast5_s2.err
run merged2_voxel6 --workload merged2 --steps 30 --warmup 5
for f in $O/bench_fast5_s3.json $O/bench_fast5_s2.json $O/bench_merged2_voxel6.json; do python - $f <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1])); r=d['roofline']
    print(sys.argv[1], 'fps %.1f e2e %.1f mups %.0f frac %.3f'%(d['value'], d['e2e']['value'], d['mvoxel_updates_per_s'], r['frac']), {k:round(v,4) for k,v in r['phase_ms_per_frame'].items()})
    print('  timeline', r.get('solve_kernel_timeline_last_profiled_frame'))
except Exception as e:
    print(sys.argv[1], 'ERR', e); print(open(sys.argv[1].replace('.json','.err')).read()[-1500:])
PY
done
