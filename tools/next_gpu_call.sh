#!/bin/bash
# GPU call 17: final verification of the round (run from the repo root under gpurun; everything lands in gpurun_out/r02/)
set -u
O=gpurun_out/r02
mkdir -p $O
timeout 1500 python -m pytest tests -q -m gpu -x 2>&1 | tail -12 > $O/gpu_suite_17.log
tail -5 $O/gpu_suite_17.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke_17.log 2>&1; tail -3 $O/smoke_17.log
q() { env "$@" timeout 300 python bench.py --quick --workload $WL --steps $ST --warmup 5 2>/dev/null | python -c 'import json,sys; d=json.loads(sys.stdin.read()); print("%.1f fps %.4f ms" % (d["value"], d["ms_per_step"]))'; }
WL=merged2; ST=30
for v in "KSG_NONE=1" "KSG_DEEP_HOT=0"; do echo "== merged2 $v: $(q $v)" | tee -a $O/tuning_17.log; done
WL=fast5; ST=100
for v in "KSG_NONE=1"; do echo "== fast5 $v: $(q $v)" | tee -a $O/tuning_17.log; done
timeout 1500 python bench.py --steps 100 --warmup 10 > $O/bench_final_17.json 2> $O/bench_final_17.err
python - $O/bench_final_17.json <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1]))
    def show(name, d):
        r=d['roofline']
        print(name, 'fps %.1f e2e %.1f (sync %.1f) mups %.0f frac %.3f frame_frac %.4f traffic %s'%(d['value'], d['e2e']['value'], d['e2e']['sync_value'], d['mvoxel_updates_per_s'], r['frac'], r['frame_frac'], r['traffic']), {k:round(v,4) for k,v in r['phase_ms_per_frame'].items()})
        print('  shim', json.dumps(d.get('e2e_shim'))[:300])
        print('  cpu', json.dumps(d.get('cpu_baseline'))[:200], 'clocks', d.get('clocks'), 'launches', d.get('gpu_launches'))
    show('fast5', d)
    for k,v in d['workloads'].items(): show(k, v)
except Exception as e:
    print('ERR', e); print(open(sys.argv[1].replace('.json','.err')).read()[-2500:])
PY
timeout 600 python bench.py --impl reference --steps 5 --warmup 3 > $O/bench_fast5_reference_17.json 2>/dev/null
cut -c1-200 $O/bench_fast5_reference_17.json
NCU="ncu --clock-control none"
timeout 600 $NCU --metrics gpu__time_duration.sum -c 400 --csv --log-file $O/launches_fast5_17.csv python tools/run_frames.py fast5 12 > /dev/null 2>&1
timeout 600 $NCU --metrics gpu__time_duration.sum -c 400 --csv --log-file $O/launches_merged2_17.csv python tools/run_frames.py merged2 6 > /dev/null 2>&1
timeout 900 $NCU --set full --import-source on -k regex:k_voxel_apply -s 9 -c 3 -o $O/prof_apply_merged2_17 -f python tools/run_frames.py merged2 5 > $O/ncu_apply2_17.log 2>&1
if [ -f $O/prof_apply_merged2_17.ncu-rep ]; then
  ncu -i $O/prof_apply_merged2_17.ncu-rep --page raw --csv > $O/prof_apply_merged2_17.raw.csv 2>/dev/null
  ncu -i $O/prof_apply_merged2_17.ncu-rep --page details > $O/prof_apply_merged2_17.details.txt 2>/dev/null
fi
ls -la $O | tail -6
