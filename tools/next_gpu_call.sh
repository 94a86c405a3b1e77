#!/bin/bash
# GPU call 12 (run from the repo root under gpurun; everything lands in gpurun_out/r02/)
set -u
O=gpurun_out/r02
mkdir -p $O
T="tests/test_gpu_parity.py tests/test_gpu_ref_golden.py tests/test_gpu_more.py tests/test_gpu_mesh.py"
timeout 1200 python -m pytest $T -q -m gpu -x 2>&1 | tail -15 > $O/gpu_quick_12.log
tail -6 $O/gpu_quick_12.log
q() { env "$@" timeout 300 python bench.py --quick --workload $WL --steps $ST --warmup 5 2>/dev/null | python -c 'import json,sys; d=json.loads(sys.stdin.read()); print("%.1f fps %.4f ms" % (d["value"], d["ms_per_step"]))'; }
WL=merged2; ST=30
for v in "KSG_NONE=1" "KSG_SHORT_THREAD=0" "KSG_SHORT_T_CTAS=1" "KSG_SHORT_T_CTAS=3" "KSG_SHORT_T_CTAS=4" "KSG_LONG_THREADS=128 KSG_LONG_GRID=296" "KSG_LONG_THREADS=128 KSG_LONG_GRID=296 KSG_SHORT_T_CTAS=3" "KSG_LONG_LEN=96" "KSG_LONG_LEN=512" "KSG_LONG_LEN=1024" "KSG_LONG_LEN=4096" "KSG_LONG_LEN=1024 KSG_SHORT_T_CTAS=3" "KSG_LONG_LEN=1024 KSG_LONG_GRID=296"; do
  echo "== merged2 $v: $(q $v)" | tee -a $O/tuning_12.log
done
WL=fast5; ST=100
for v in "KSG_NONE=1" "KSG_GROUP0=16384" "KSG_GROUP0=8192 KSG_GROUP_MUL=16" "KSG_SOLVE_THREADS=512" "KSG_GROUP0=512"; do
  echo "== fast5 $v: $(q $v)" | tee -a $O/tuning_12.log
done
# two solve kernels side by side? (512-thread CTAs leave room for a second cooperative kernel)
for v in "KSG_NONE=1" "KSG_SOLVE_THREADS=512" "KSG_SOLVE_THREADS=256"; do
  for k in 2 4; do
    echo "== fast5 multi-sequence x$k $v: $(env $v timeout 300 python bench.py --no-cpu-baseline --shim-e2e 0 --extra-workloads '' --sequences-per-gpu $k --steps 60 --warmup 5 2>/dev/null | python -c 'import json,sys; d=json.loads(sys.stdin.read()); print("value %.1f multi %.1f" % (d["value"], d["multi_sequence"]["value"]))')" | tee -a $O/tuning_12.log
  done
done
timeout 300 python tools/diag_frames_mode.py fast5 > $O/diag_frames_mode.txt 2>&1; tail -22 $O/diag_frames_mode.txt
timeout 1200 python bench.py --no-cpu-baseline --steps 100 --warmup 10 > $O/bench_full_12.json 2> $O/bench_full_12.err
python - $O/bench_full_12.json <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1]))
    def show(name, d):
        r=d['roofline']
        print(name, 'fps %.1f e2e %.1f (sync %.1f) mups %.0f frac %.3f frame_frac %.4f'%(d['value'], d['e2e']['value'], d['e2e']['sync_value'], d['mvoxel_updates_per_s'], r['frac'], r['frame_frac']), {k:round(v,4) for k,v in r['phase_ms_per_frame'].items()})
        print('  multi', d.get('multi_sequence'))
        print('  timeline', json.dumps(r.get('solve_kernel_timeline_last_profiled_frame')))
    show('fast5', d)
    for k,v in d['workloads'].items(): show(k, v)
except Exception as e:
    print('ERR', e); print(open(sys.argv[1].replace('.json','.err')).read()[-2500:])
PY
NCU="ncu --clock-control none"
timeout 900 $NCU --set full --import-source on -k regex:k_voxel_apply -s 6 -c 2 -o $O/prof_apply_merged2_12 -f python tools/run_frames.py merged2 5 > $O/ncu_apply2_12.log 2>&1
timeout 900 $NCU --set full -k regex:Onesweep -s 44 -c 4 -o $O/prof_sort_merged2_12 -f python tools/run_frames.py merged2 5 > $O/ncu_sort2_12.log 2>&1
for f in apply_merged2_12 sort_merged2_12; do
  if [ -f $O/prof_$f.ncu-rep ]; then
    ncu -i $O/prof_$f.ncu-rep --page raw --csv > $O/prof_$f.raw.csv 2>/dev/null
    ncu -i $O/prof_$f.ncu-rep --page details > $O/prof_$f.details.txt 2>/dev/null
  fi
done
ls -la $O | tail -12
