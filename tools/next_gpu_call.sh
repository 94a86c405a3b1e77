#!/bin/bash
# GPU call 19: unrolled deep pipelines for the hot voxels
set -u
O=gpurun_out/r02
mkdir -p $O
T="tests/test_gpu_parity.py tests/test_gpu_ref_golden.py tests/test_gpu_more.py tests/test_gpu_delta_merge.py"
timeout 900 python -m pytest $T -q -m gpu -x 2>&1 | tail -6 > $O/gpu_quick_19.log
tail -4 $O/gpu_quick_19.log
q() { env "$@" timeout 300 python bench.py --quick --workload $WL --steps $ST --warmup 5 2>/dev/null | python -c 'import json,sys; d=json.loads(sys.stdin.read()); print("%.1f fps %.4f ms" % (d["value"], d["ms_per_step"]))'; }
WL=merged2; ST=30
for v in "KSG_NONE=1" "KSG_SHORT_T_CTAS=3" "KSG_SHORT_T_CTAS=1" "KSG_LONG_SERIAL=0" "KSG_DEEP_HOT=0"; do
  echo "== merged2 $v: $(q $v)" | tee -a $O/tuning_19.log
done
WL=fast5; ST=100
echo "== fast5: $(q KSG_NONE=1)" | tee -a $O/tuning_19.log
NCU="ncu --clock-control none"
timeout 900 $NCU --set full --import-source on -k regex:k_voxel_apply -s 9 -c 3 -o $O/prof_apply_merged2_19 -f python tools/run_frames.py merged2 5 > $O/ncu_apply2_19.log 2>&1
if [ -f $O/prof_apply_merged2_19.ncu-rep ]; then
  ncu -i $O/prof_apply_merged2_19.ncu-rep --page raw --csv > $O/prof_apply_merged2_19.raw.csv 2>/dev/null
  ncu -i $O/prof_apply_merged2_19.ncu-rep --page details > $O/prof_apply_merged2_19.details.txt 2>/dev/null
  grep -E "^  [a-zA-Z_:<>, ()0-9*&]+\(|    Duration" $O/prof_apply_merged2_19.details.txt | cut -c1-100
fi
