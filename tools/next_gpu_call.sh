#!/bin/bash
# GPU call 16: deep-pipeline instance for the hot voxels of `merged`
set -u
O=gpurun_out/r02
mkdir -p $O
T="tests/test_gpu_parity.py tests/test_gpu_ref_golden.py tests/test_gpu_more.py"
timeout 900 python -m pytest $T -q -m gpu -x -k "merged or Merged or MERGED" 2>&1 | tail -6 > $O/gpu_quick_16.log
tail -4 $O/gpu_quick_16.log
q() { env "$@" timeout 300 python bench.py --quick --workload $WL --steps $ST --warmup 5 2>/dev/null | python -c 'import json,sys; d=json.loads(sys.stdin.read()); print("%.1f fps %.4f ms" % (d["value"], d["ms_per_step"]))'; }
WL=merged2; ST=30
for v in "KSG_NONE=1" "KSG_DEEP_HOT=0" "KSG_SHORT_T_CTAS=2" "KSG_LONG_THREADS=128 KSG_LONG_GRID=296" "KSG_LONG_GRID=296"; do
  echo "== merged2 $v: $(q $v)" | tee -a $O/tuning_16.log
done
timeout 600 python bench.py --workload merged2 --steps 30 --warmup 5 --no-cpu-baseline --extra-workloads '' --shim-e2e 0 > $O/bench_merged2_16.json 2>/dev/null
python -c "
import json; d=json.load(open('$O/bench_merged2_16.json')); print('merged2 fps %.1f e2e %.1f'%(d['value'], d['e2e']['value']), {k:round(v,3) for k,v in d['roofline']['phase_ms_per_frame'].items()})"
