#!/bin/bash
# GPU call 18: balance of the three merged update kernels
set -u
O=gpurun_out/r02
mkdir -p $O
q() { env "$@" timeout 300 python bench.py --quick --workload $WL --steps $ST --warmup 5 2>/dev/null | python -c 'import json,sys; d=json.loads(sys.stdin.read()); print("%.1f fps %.4f ms" % (d["value"], d["ms_per_step"]))'; }
WL=merged2; ST=30
for v in "KSG_NONE=1" "KSG_SHORT_T_CTAS=2" "KSG_SHORT_T_CTAS=3" "KSG_LONG_SERIAL=1 KSG_SHORT_T_CTAS=2" "KSG_LONG_SERIAL=1 KSG_SHORT_T_CTAS=3" "KSG_LONG_SERIAL=1 KSG_SHORT_T_CTAS=4" "KSG_LONG_SERIAL=1 KSG_SHORT_T_CTAS=3 KSG_LONG_GRID=296" "KSG_LONG_SERIAL=1 KSG_SHORT_T_CTAS=3 KSG_LONG_LEN=512" "KSG_LONG_SERIAL=1 KSG_SHORT_T_CTAS=3 KSG_LONG_LEN=128"; do
  echo "== merged2 $v: $(q $v)" | tee -a $O/tuning_18.log
done
