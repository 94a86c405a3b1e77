#!/bin/bash
# GPU call 21: final verification of the round (run from the repo root under gpurun; everything lands in gpurun_out/r02/)
set -u
O=gpurun_out/r02
mkdir -p $O
q() { env "$@" timeout 300 python bench.py --quick --workload $WL --steps $ST --warmup 5 2>/dev/null | python -c 'import json,sys; d=json.loads(sys.stdin.read()); print("%.1f fps %.4f ms" % (d["value"], d["ms_per_step"]))'; }
WL=merged2; ST=30
for v in "KSG_NONE=1" "KSG_DEEP_HOT=0"; do echo "== merged2 $v: $(q $v)" | tee -a $O/tuning_21.log; done
timeout 1500 python -m pytest tests -q -m gpu -x 2>&1 | tail -12 > $O/gpu_suite_21.log
tail -5 $O/gpu_suite_21.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke_21.log 2>&1; tail -3 $O/smoke_21.log
timeout 1500 python bench.py --steps 100 --warmup 10 > $O/bench_final_21.json 2> $O/bench_final_21.err
python - $O/bench_final_21.json <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1]))
    def show(name, d):
        r=d['roofline']
        print(name, 'fps %.1f e2e %.1f (sync %.1f) mups %.0f frac %.3f frame_frac %.4f traffic %s'%(d['value'], d['e2e']['value'], d['e2e']['sync_value'], d['mvoxel_updates_per_s'], r['frac'], r['frame_frac'], r['traffic']), {k:round(v,4) for k,v in r['phase_ms_per_frame'].items()})
        print('  cpu', json.dumps(d.get('cpu_baseline'))[:160], 'clocks', d.get('clocks'), 'launches', d.get('gpu_launches'))
    show('fast5', d)
    for k,v in d['workloads'].items(): show(k, v)
except Exception as e:
    print('ERR', e); print(open(sys.argv[1].replace('.json','.err')).read()[-2500:])
PY
timeout 600 python bench.py --impl reference --steps 5 --warmup 3 > $O/bench_fast5_reference_21.json 2>/dev/null
cut -c1-160 $O/bench_fast5_reference_21.json
NCU="ncu --clock-control none"
timeout 600 $NCU --metrics gpu__time_duration.sum -c 400 --csv --log-file $O/launches_merged2_21.csv python tools/run_frames.py merged2 6 > /dev/null 2>&1
timeout 900 $NCU --set full --import-source on -k regex:k_voxel_apply -s 9 -c 3 -o $O/prof_apply_merged2_21 -f python tools/run_frames.py merged2 5 > $O/ncu_apply2_21.log 2>&1
if [ -f $O/prof_apply_merged2_21.ncu-rep ]; then
  ncu -i $O/prof_apply_merged2_21.ncu-rep --page raw --csv > $O/prof_apply_merged2_21.raw.csv 2>/dev/null
  ncu -i $O/prof_apply_merged2_21.ncu-rep --page details > $O/prof_apply_merged2_21.details.txt 2>/dev/null
  grep -E "^  [a-zA-Z_:<>, ()0-9*&]+\(|    Duration" $O/prof_apply_merged2_21.details.txt | cut -c1-100
fi
