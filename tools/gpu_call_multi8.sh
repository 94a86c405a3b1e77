#!/bin/bash
# 8-GPU call (charged 8x): gpurun --gpus 8 -- 'bash tools/gpu_call_multi8.sh'
set -u
N=8
O=gpurun_out/r02
mkdir -p $O
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
timeout 600 python -m pytest tests/test_gpu_multi.py -q -m gpu -x 2>&1 | tail -6 > $O/gpu_multi_n$N.log
tail -3 $O/gpu_multi_n$N.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611"
# the driver's own command line at N = 8: replicas headline + the two shared-map modes inside `workloads`
timeout 900 $TR bench.py --gpus $N --steps 50 --warmup 5 > $O/bench_seq_fast5_n$N.json 2> $O/bench_seq_fast5_n$N.err
# BASELINE configs[3]: 1280x720, 5 cm, C = 150, batches of 8 frames, one frame per GPU
timeout 900 $TR bench.py --gpus $N --steps 15 --warmup 3 --workload fast5_720p_c150 --sharding frames > $O/bench_frames_720p_c150_n$N.json 2> $O/bench_frames_720p_c150_n$N.err
# BASELINE configs[4]: 3840x2160, 1 cm, C = 40, merged, one map spatially sharded over the 8 GPUs
timeout 1500 $TR bench.py --gpus $N --steps 2 --warmup 3 --workload merged1_4k_c40 --sharding spatial --no-cpu-baseline --profile-frames 2 > $O/bench_spatial_merged1_4k_c40_n$N.json 2> $O/bench_spatial_merged1_4k_c40_n$N.err
for f in bench_seq_fast5_n$N bench_frames_720p_c150_n$N bench_spatial_merged1_4k_c40_n$N; do
python - $O/$f <<'PY'
import json,sys
p=sys.argv[1]
try:
    txt=open(p+'.json').read()
    line=[l for l in txt.splitlines() if l.startswith('{')][-1]
    d=json.loads(line)
    print(p, 'stdout lines', len(txt.splitlines()), 'fps %.2f e2e %.2f scaling %s'%(d['value'], d['e2e']['value'], d['scaling']), (d.get('collective') or {}).get('bytes_per_step'), d.get('phase_ms_per_step'), (d.get('roofline') or {}).get('phase_ms_per_frame'))
    for k,v in (d.get('workloads') or {}).items(): print('   wl', k, (v or {}).get('value'), (v or {}).get('error'), (v or {}).get('phase_ms_per_step'), ((v or {}).get('roofline') or {}).get('phase_ms_per_frame'))
except Exception as e:
    print(p, 'ERR', e); print(open(p+'.err').read()[-1500:])
PY
done
