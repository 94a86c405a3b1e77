#!/bin/bash
# GPU call 20: size of the hot-voxel instance
set -u
O=gpurun_out/r02
mkdir -p $O
q() { env "$@" timeout 300 python bench.py --quick --workload $WL --steps $ST --warmup 5 2>/dev/null | python -c 'import json,sys; d=json.loads(sys.stdin.read()); print("%.1f fps %.4f ms" % (d["value"], d["ms_per_step"]))'; }
WL=merged2; ST=30
for v in "KSG_DEEP_THREADS=64" "KSG_DEEP_THREADS=128" "KSG_DEEP_THREADS=256" "KSG_DEEP_THREADS=256 KSG_SHORT_T_CTAS=1" "KSG_DEEP_THREADS=128 KSG_SHORT_T_CTAS=1" "KSG_DEEP_THREADS=256 KSG_LONG_SERIAL=0"; do
  echo "== merged2 $v: $(q $v)" | tee -a $O/tuning_20.log
done
