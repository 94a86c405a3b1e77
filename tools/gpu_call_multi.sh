#!/bin/bash
# multi-GPU call: gpurun --gpus N -- 'bash tools/gpu_call_multi.sh N'
set -u
N=${1:-2}
O=gpurun_out/r02
mkdir -p $O
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
timeout 900 python -m pytest tests/test_gpu_multi.py tests/test_gpu_delta_merge.py -q -m gpu -x 2>&1 | tail -8 > $O/gpu_multi_n$N.log
tail -4 $O/gpu_multi_n$N.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611"
timeout 900 $TR bench.py --gpus $N --steps 50 --warmup 5 --no-cpu-baseline > $O/bench_seq_fast5_n$N.json 2> $O/bench_seq_fast5_n$N.err
timeout 900 $TR bench.py --gpus $N --steps 20 --warmup 5 --workload merged2 --sharding spatial --no-cpu-baseline > $O/bench_spatial_merged2_n$N.json 2> $O/bench_spatial_merged2_n$N.err
timeout 900 $TR bench.py --gpus $N --steps 20 --warmup 3 --workload fast5_720p_c150 --sharding frames > $O/bench_frames_720p_c150_n$N.json 2> $O/bench_frames_720p_c150_n$N.err
timeout 900 $TR bench.py --gpus $N --steps 30 --warmup 3 --workload fast5 --sharding frames > $O/bench_frames_fast5_n$N.json 2> $O/bench_frames_fast5_n$N.err
KSG_FRAMES_BY_BLOCKS=1 timeout 900 $TR bench.py --gpus $N --steps 10 --warmup 3 --workload fast5 --sharding frames > $O/bench_frames_fast5_blocks_n$N.json 2> $O/bench_frames_fast5_blocks_n$N.err
timeout 1500 $TR bench.py --gpus $N --steps 3 --warmup 3 --workload merged1_4k_c40 --sharding spatial --no-cpu-baseline --profile-frames 2 > $O/bench_spatial_merged1_4k_c40_n$N.json 2> $O/bench_spatial_merged1_4k_c40_n$N.err
for f in bench_seq_fast5_n$N bench_spatial_merged2_n$N bench_frames_720p_c150_n$N bench_frames_fast5_n$N bench_frames_fast5_blocks_n$N bench_spatial_merged1_4k_c40_n$N; do
python - $O/$f <<'PY'
import json,sys
p=sys.argv[1]
try:
    line=[l for l in open(p+'.json').read().splitlines() if l.startswith('{')][-1]
    d=json.loads(line)
    print(p, 'fps %.2f e2e %.2f scaling %s'%(d['value'], d['e2e']['value'], d['scaling']), (d.get('collective') or {}).get('bytes_per_step'), d.get('phase_ms_per_step'), (d.get('roofline') or {}).get('phase_ms_per_frame'))
    for k,v in (d.get('workloads') or {}).items(): print('   wl', k, (v or {}).get('value'), (v or {}).get('error'), (v or {}).get('phase_ms_per_step'))
except Exception as e:
    print(p, 'ERR', e); print(open(p+'.err').read()[-1500:])
PY
done
