#!/bin/bash
# GPU call 22: L2 persistence window for the (L * freq) rows
set -u
O=gpurun_out/r02
mkdir -p $O
q() { env "$@" timeout 300 python bench.py --quick --workload $WL --steps $ST --warmup 5 2>/dev/null | python -c 'import json,sys; d=json.loads(sys.stdin.read()); print("%.1f fps %.4f ms" % (d["value"], d["ms_per_step"]))'; }
WL=merged2; ST=30
for v in "KSG_NONE=1" "KSG_L2_PERSIST=1" "KSG_NONE=2" "KSG_L2_PERSIST=1 KSG_SHORT_T_CTAS=3"; do echo "== merged2 $v: $(q $v)" | tee -a $O/tuning_22.log; done
