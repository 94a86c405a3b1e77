#!/bin/bash
# GPU call (run from the repo root under gpurun; everything lands in gpurun_out/r02/)
set -u
O=gpurun_out/r02
mkdir -p $O
timeout 180 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke_s31.log 2>&1; echo "smoke rc=$?" >> $O/smoke_s31.log
tail -4 $O/smoke_s31.log
timeout 1500 python -m pytest tests -q -m gpu -x 2>&1 | tail -15 > $O/gpu_suite_s31.log
tail -8 $O/gpu_suite_s31.log
run() { name=$1; shift; timeout 400 python bench.py --no-cpu-baseline "$@" > $O/bench_${name}.json 2> $O/bench_${name}.err; }
run fast5_s31 --steps 100 --warmup 10
run merged2_hot --workload merged2 --steps 30 --warmup 5
run merged5_hot --workload merged5 --steps 30 --warmup 5
for f in $O/bench_fast5_s31.json $O/bench_merged2_hot.json $O/bench_merged5_hot.json; do python - $f <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1])); r=d['roofline']
    print(sys.argv[1], 'fps %.1f e2e %.1f mups %.0f frac %.3f'%(d['value'], d['e2e']['value'], d['mvoxel_updates_per_s'], r['frac']), {k:round(v,4) for k,v in r['phase_ms_per_frame'].items()})
    print('  timeline', r.get('solve_kernel_timeline_last_profiled_frame'))
except Exception as e:
    print(sys.argv[1], 'ERR', e); print(open(sys.argv[1].replace('.json','.err')).read()[-1500:])
PY
done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_fast5_s31.csv python tools/run_frames.py fast5 14 > $O/ncu_fast5.log 2>&1
