#!/bin/bash
# GPU call 13 (run from the repo root under gpurun; everything lands in gpurun_out/r02/)
set -u
O=gpurun_out/r02
mkdir -p $O
timeout 1500 python -m pytest tests -q -m gpu -x 2>&1 | tail -12 > $O/gpu_suite_13.log
tail -5 $O/gpu_suite_13.log
q() { env "$@" timeout 300 python bench.py --quick --workload $WL --steps $ST --warmup 5 2>/dev/null | python -c 'import json,sys; d=json.loads(sys.stdin.read()); print("%.1f fps %.4f ms" % (d["value"], d["ms_per_step"]))'; }
WL=merged2; ST=30
for v in "KSG_NONE=1" "KSG_LONG_GRID=296" "KSG_LONG_THREADS=128 KSG_LONG_GRID=444" "KSG_LONG_THREADS=128 KSG_LONG_GRID=296" "KSG_LONG_GRID=222"; do
  echo "== merged2 $v: $(q $v)" | tee -a $O/tuning_13.log
done
timeout 1500 python bench.py --steps 100 --warmup 10 > $O/bench_final_13.json 2> $O/bench_final_13.err
python - $O/bench_final_13.json <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1]))
    def show(name, d):
        r=d['roofline']
        print(name, 'fps %.1f e2e %.1f (sync %.1f) mups %.0f frac %.3f frame_frac %.4f traffic %s'%(d['value'], d['e2e']['value'], d['e2e']['sync_value'], d['mvoxel_updates_per_s'], r['frac'], r['frame_frac'], r['traffic']), {k:round(v,4) for k,v in r['phase_ms_per_frame'].items()})
        print('  multi', d.get('multi_sequence'))
        print('  shim', json.dumps(d.get('e2e_shim'))[:600])
        print('  cpu', json.dumps(d.get('cpu_baseline'))[:400])
        print('  clocks', d.get('clocks'), 'launches', d.get('gpu_launches'))
    show('fast5', d)
    for k,v in d['workloads'].items(): show(k, v)
except Exception as e:
    print('ERR', e); print(open(sys.argv[1].replace('.json','.err')).read()[-2500:])
PY
timeout 600 python bench.py --impl reference --steps 5 --warmup 3 > $O/bench_fast5_reference_13.json 2>/dev/null
cut -c1-300 $O/bench_fast5_reference_13.json
timeout 900 python bench.py --workload fast5_720p_c150 --steps 30 --warmup 5 --no-cpu-baseline --extra-workloads '' --shim-e2e 0 --sequences-per-gpu 0 > $O/bench_fast5_720p_c150_n1.json 2> $O/bench_fast5_720p_c150_n1.err
timeout 1200 python bench.py --workload merged1_4k_c40 --steps 3 --warmup 3 --no-cpu-baseline --extra-workloads '' --shim-e2e 0 --profile-frames 2 > $O/bench_merged1_4k_c40_n1.json 2> $O/bench_merged1_4k_c40_n1.err
for f in bench_fast5_720p_c150_n1 bench_merged1_4k_c40_n1; do
python - $O/$f.json <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1])); r=d['roofline']
    print(sys.argv[1], 'fps %.2f e2e %.2f mups %.0f frame_frac %.4f blocks %s'%(d['value'], d['e2e']['value'], d['mvoxel_updates_per_s'], r['frame_frac'], d['config'].get('map_blocks_after_run')), {k:round(v,4) for k,v in r['phase_ms_per_frame'].items()})
except Exception as e:
    print(sys.argv[1], 'ERR', e); print(open(sys.argv[1].replace('.json','.err')).read()[-1500:])
PY
done
NCU="ncu --clock-control none"
timeout 600 $NCU --metrics gpu__time_duration.sum -c 400 --csv --log-file $O/launches_fast5_13.csv python tools/run_frames.py fast5 12 > /dev/null 2>&1
timeout 600 $NCU --metrics gpu__time_duration.sum -c 400 --csv --log-file $O/launches_merged2_13.csv python tools/run_frames.py merged2 6 > /dev/null 2>&1
ls -la $O | tail -8
