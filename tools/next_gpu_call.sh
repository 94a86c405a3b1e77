#!/bin/bash
# GPU call 10 (run from the repo root under gpurun; everything lands in gpurun_out/r02/)
set -u
O=gpurun_out/r02
mkdir -p $O
timeout 1500 python -m pytest tests -q -m gpu -x 2>&1 | tail -12 > $O/gpu_suite_10.log
tail -5 $O/gpu_suite_10.log
timeout 1200 python bench.py --steps 100 --warmup 10 > $O/bench_full_10.json 2> $O/bench_full_10.err
python - $O/bench_full_10.json <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1]))
    def show(name, d):
        r=d['roofline']
        print(name, 'fps %.1f e2e %.1f (sync %.1f) mups %.0f frac %.3f frame_frac %.4f'%(d['value'], d['e2e']['value'], d['e2e']['sync_value'], d['mvoxel_updates_per_s'], r['frac'], r['frame_frac']), {k:round(v,4) for k,v in r['phase_ms_per_frame'].items()})
        print('  multi', d.get('multi_sequence'))
        print('  shim', json.dumps(d.get('e2e_shim')))
        print('  cpu', json.dumps(d.get('cpu_baseline'))[:300])
        print('  timeline', json.dumps(r.get('solve_kernel_timeline_last_profiled_frame')))
    show('fast5', d)
    for k,v in d['workloads'].items(): show(k, v)
except Exception as e:
    print('ERR', e); print(open(sys.argv[1].replace('.json','.err')).read()[-2500:])
PY
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > $O/bench_fast5_reference_10.json 2>/dev/null
cut -c1-400 $O/bench_fast5_reference_10.json
# tuning sweep of the solve kernel (quick legs only)
for v in "KSG_GROUP0=1024" "KSG_GROUP0=512" "KSG_GROUP0=2048" "KSG_GROUP0=8192" "KSG_GROUP0=1000000" "KSG_GROUP0=512 KSG_GROUP_MUL=8" "KSG_GROUP0=128 KSG_GROUP_MUL=2" "KSG_SOLVE_THREADS=512" "KSG_SOLVE_THREADS=256" "KSG_SOLVE_THREADS=512 KSG_SOLVE_CTAS_PER_SM=1"; do
  echo "== $v: $(env $v timeout 300 python bench.py --quick --steps 100 --warmup 10 2>/dev/null | python -c 'import json,sys; d=json.loads(sys.stdin.read()); print("%.1f fps %.4f ms" % (d["value"], d["ms_per_step"]))')" | tee -a $O/tuning_10.log
done
# merged2 variants of the voxel apply
for v in "KSG_NONE=1" "KSG_SHORT_CTAS=3" "KSG_SHORT_CTAS=5 KSG_LONG_THREADS=64 KSG_LONG_GRID=296" "KSG_SHORT_CTAS=5 KSG_LONG_THREADS=64 KSG_LONG_GRID=592" "KSG_SHORT_CTAS=4 KSG_LONG_THREADS=128 KSG_LONG_GRID=296" "KSG_SHORT_CTAS=4 KSG_LONG_THREADS=128 KSG_LONG_GRID=148" "KSG_LONG_GRID=16" "KSG_EMIT_WARP=1" "KSG_HOT_KERNEL=1"; do
  echo "== $v: $(env $v timeout 300 python bench.py --quick --workload merged2 --steps 30 --warmup 5 2>/dev/null | python -c 'import json,sys; d=json.loads(sys.stdin.read()); print("%.1f fps %.4f ms" % (d["value"], d["ms_per_step"]))')" | tee -a $O/tuning_10.log
done
# ncu: launch lists (cold, serialised: shares only) and one --set full capture per dominant kernel
NCU="ncu --clock-control none"
timeout 600 $NCU --metrics gpu__time_duration.sum -c 400 --csv --log-file $O/launches_fast5_10.csv python tools/run_frames.py fast5 12 > /dev/null 2>&1
timeout 600 $NCU --metrics gpu__time_duration.sum -c 400 --csv --log-file $O/launches_merged2_10.csv python tools/run_frames.py merged2 6 > /dev/null 2>&1
timeout 900 $NCU --set full --import-source on -k regex:k_fast_solve3 -s 6 -c 1 -o $O/prof_solve3_fast5 -f python tools/run_frames.py fast5 8 > $O/ncu_solve3.log 2>&1
timeout 600 $NCU --set full --import-source on -k regex:k_tile_apply_fast -s 6 -c 1 -o $O/prof_apply_fast5 -f python tools/run_frames.py fast5 8 > $O/ncu_apply5.log 2>&1
timeout 900 $NCU --set full --import-source on -k regex:k_voxel_apply -s 6 -c 2 -o $O/prof_apply_merged2 -f python tools/run_frames.py merged2 5 > $O/ncu_apply2.log 2>&1
timeout 900 $NCU --set full -k regex:DeviceRadixSort -s 24 -c 8 -o $O/prof_sort_merged2 -f python tools/run_frames.py merged2 5 > $O/ncu_sort2.log 2>&1
for f in solve3_fast5 apply_fast5 apply_merged2 sort_merged2; do
  if [ -f $O/prof_$f.ncu-rep ]; then
    ncu -i $O/prof_$f.ncu-rep --page raw --csv > $O/prof_$f.raw.csv 2>/dev/null
    ncu -i $O/prof_$f.ncu-rep --page details > $O/prof_$f.details.txt 2>/dev/null
  fi
done
ls -la $O | tail -30
