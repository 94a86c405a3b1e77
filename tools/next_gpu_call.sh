#!/bin/bash
# GPU call (run from the repo root under gpurun; everything lands in gpurun_out/r02/)
set -u
O=gpurun_out/r02
mkdir -p $O
T="tests/test_gpu_parity.py tests/test_gpu_ref_golden.py"
timeout 900 python -m pytest $T tests/test_gpu_more.py -q -m gpu -x 2>&1 | tail -8 > $O/gpu_quick_9.log
tail -4 $O/gpu_quick_9.log
if ! grep -q " passed" $O/gpu_quick_9.log || grep -q "failed" $O/gpu_quick_9.log; then
  # which of the new pieces breaks parity?
  for v in "KSG_EMIT_WARP=0" "KSG_HOT_KERNEL=0" "KSG_EMIT_WARP=0 KSG_HOT_KERNEL=0" "KSG_NO_UPDATE_LOG=1"; do
    echo "== $v" >> $O/gpu_bisect_9.log
    env $v timeout 600 python -m pytest $T tests/test_gpu_more.py -q -m gpu -x 2>&1 | tail -4 >> $O/gpu_bisect_9.log
  done
  cat $O/gpu_bisect_9.log
fi
timeout 1200 python bench.py --no-cpu-baseline --steps 100 --warmup 10 > $O/bench_full_9.json 2> $O/bench_full_9.err
python - $O/bench_full_9.json <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1]))
    def show(name, d):
        r=d['roofline']
        print(name, 'fps %.1f e2e %.1f (sync %.1f) mups %.0f frac %.3f frame_frac %.4f'%(d['value'], d['e2e']['value'], d['e2e']['sync_value'], d['mvoxel_updates_per_s'], r['frac'], r['frame_frac']), {k:round(v,4) for k,v in r['phase_ms_per_frame'].items()})
        print('  multi', d.get('multi_sequence'))
        print('  shim', json.dumps(d.get('e2e_shim')))
        print('  timeline', json.dumps(r.get('solve_kernel_timeline_last_profiled_frame')))
    show('fast5', d)
    for k,v in d['workloads'].items(): show(k, v)
except Exception as e:
    print('ERR', e); print(open(sys.argv[1].replace('.json','.err')).read()[-2500:])
PY
KSG_HOT_KERNEL=0 KSG_EMIT_WARP=0 timeout 600 python bench.py --no-cpu-baseline --workload merged2 --steps 30 --warmup 5 --extra-workloads "" --shim-e2e 0 > $O/bench_merged2_nohotk.json 2>/dev/null
python -c "
import json; d=json.load(open('$O/bench_merged2_nohotk.json')); print('merged2 without hot kernel / warp emit: fps %.1f'%d['value'], {k:round(v,3) for k,v in d['roofline']['phase_ms_per_frame'].items()})"
# tuning sweep of the solve kernel (quick legs only)
for v in "KSG_GROUP0=512" "KSG_GROUP0=2048" "KSG_GROUP0=8192" "KSG_GROUP0=1000000" "KSG_GROUP0=512 KSG_GROUP_MUL=8" "KSG_GROUP0=128 KSG_GROUP_MUL=2" "KSG_SOLVE_THREADS=512" "KSG_SOLVE_THREADS=256" "KSG_SOLVE_THREADS=512 KSG_SOLVE_CTAS_PER_SM=1"; do
  echo "== $v: $(env $v timeout 300 python bench.py --quick --steps 100 --warmup 10 2>/dev/null | python -c 'import json,sys; d=json.loads(sys.stdin.read()); print("%.1f fps %.4f ms" % (d["value"], d["ms_per_step"]))')" | tee -a $O/tuning_9.log
done
