#!/bin/bash
# 2 GPUs: NCCL product test + bench (device-resident and e2e with slab learner / column-group actor under torchrun)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
N=2
echo "== pytest multi"; timeout -s KILL 400 python -m pytest tests/test_gpu_multi.py -q 2>&1 | tail -5
echo "== bench n=$N"; timeout -s KILL 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 12 --warmup 4 > gpurun_out/r2_bench_n${N}_c.json 2> gpurun_out/r2_bench_n${N}_c.err; tail -1 gpurun_out/r2_bench_n${N}_c.json | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); e=d['e2e']; print('value', int(d['value']), 'ms', round(d['ms_per_step'],3), 'e2e', int(e['value']), round(e['ms_per_step'],1), e['learner_thread_ms_per_step'], e['actor_last_sample_ms'], 'k1', round(d['roofline_k1']['frac'],3))"
tail -3 gpurun_out/r2_bench_n${N}_c.err
