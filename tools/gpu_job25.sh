#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
N=${NGPU:-8}
timeout -s KILL 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r2_bench_n$N.json 2> gpurun_out/r2_bench_n$N.err
tail -1 gpurun_out/r2_bench_n$N.json | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('N', d['n_gpus'], 'value', int(d['value']), 'ms', round(d['ms_per_step'],3), 'e2e', d.get('e2e',{}) and int(d['e2e']['value']))"
grep -v "OMP_NUM\|^\*\*\*" gpurun_out/r2_bench_n$N.err | tail -3
