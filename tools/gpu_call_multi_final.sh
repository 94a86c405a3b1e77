#!/bin/bash
# last multi-GPU check of the round (2 GPUs): real-NCCL parity test + the driver's own N = 2 command line with the final code
set -u
N=2
O=gpurun_out/r02
mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_multi.py -q -m gpu -x 2>&1 | tail -4 > $O/gpu_multi_final_n$N.log
tail -2 $O/gpu_multi_final_n$N.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611"
timeout 900 $TR bench.py --gpus $N --steps 100 --warmup 10 > $O/bench_seq_fast5_final_n$N.json 2> $O/bench_seq_fast5_final_n$N.err
python - $O/bench_seq_fast5_final_n$N <<'PY'
import json,sys
p=sys.argv[1]
try:
    txt=open(p+'.json').read()
    d=json.loads([l for l in txt.splitlines() if l.startswith('{')][-1])
    print(p, 'stdout lines', len(txt.splitlines()), 'fps %.1f e2e %.1f per-GPU %.1f'%(d['value'], d['e2e']['value'], d['value']/d['n_gpus']))
    for k,v in (d.get('workloads') or {}).items(): print('   wl', k, (v or {}).get('value'), (v or {}).get('error'), (v or {}).get('phase_ms_per_step'))
except Exception as e:
    print(p, 'ERR', e); print(open(p+'.err').read()[-1500:])
PY
