#!/bin/bash
# multi-GPU call: gpurun --gpus N -- 'bash tools/gpu_call_multi.sh N'
set -u
N=${1:-2}
O=gpurun_out/r02
mkdir -p $O
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
timeout 900 python -m pytest tests/test_gpu_multi.py -q -m gpu -x 2>&1 | tail -6 > $O/gpu_multi_n$N.log
tail -4 $O/gpu_multi_n$N.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611"
timeout 900 $TR bench.py --gpus $N --steps 20 --warmup 5 --workload merged2 --sharding spatial --no-cpu-baseline > $O/bench_spatial_merged2_n$N.json 2> $O/bench_spatial_merged2_n$N.err
timeout 900 $TR bench.py --gpus $N --steps 10 --warmup 3 --workload fast5_720p_c150 --sharding frames > $O/bench_frames_720p_c150_n$N.json 2> $O/bench_frames_720p_c150_n$N.err
timeout 900 $TR bench.py --gpus $N --steps 10 --warmup 3 --workload fast5 --sharding frames > $O/bench_frames_fast5_n$N.json 2> $O/bench_frames_fast5_n$N.err
timeout 900 $TR bench.py --gpus $N --steps 50 --warmup 5 --no-cpu-baseline > $O/bench_seq_fast5_n$N.json 2> $O/bench_seq_fast5_n$N.err
for f in $O/bench_spatial_merged2_n$N $O/bench_frames_720p_c150_n$N $O/bench_frames_fast5_n$N $O/bench_seq_fast5_n$N; do
python - $f.json <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1]))
    print(sys.argv[1], 'fps %.1f e2e %.1f scaling %s'%(d['value'], d['e2e']['value'], d['scaling']), d.get('collective'), d.get('phase_ms_per_step'), (d.get('roofline') or {}).get('phase_ms_per_frame'))
except Exception as e:
    print(sys.argv[1], 'ERR', e); print(open(sys.argv[1].replace('.json','.err')).read()[-1800:])
PY
done
