#!/bin/bash
# GPU call (run from the repo root under gpurun; everything lands in gpurun_out/r02/)
set -u
O=gpurun_out/r02
mkdir -p $O
timeout 1500 python -m pytest tests -q -m gpu -x 2>&1 | tail -12 > $O/gpu_suite_8.log
tail -6 $O/gpu_suite_8.log
timeout 900 python bench.py --no-cpu-baseline --steps 100 --warmup 10 --extra-workloads "" > $O/bench_fast5_8.json 2> $O/bench_fast5_8.err
python - $O/bench_fast5_8.json <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1]))
    r=d['roofline']
    print('fast5 fps %.1f e2e %.1f (sync %.1f) mups %.0f'%(d['value'], d['e2e']['value'], d['e2e']['sync_value'], d['mvoxel_updates_per_s']), {k:round(v,4) for k,v in r['phase_ms_per_frame'].items()})
    print('  multi', d.get('multi_sequence'))
    print('  timeline', json.dumps(r.get('solve_kernel_timeline_last_profiled_frame')))
except Exception as e:
    print('ERR', e); print(open(sys.argv[1].replace('.json','.err')).read()[-2500:])
PY
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_merged2_8.csv python tools/run_frames.py merged2 5 > $O/ncu_merged2.log 2>&1
