#!/bin/bash
# 2 GPUs, final tree: bench under torchrun (e2e with the bidirectional all-rank copy measurement)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
N=2
timeout -s KILL 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_n${N}_d.json 2> gpurun_out/r2_bench_n${N}_d.err; tail -1 gpurun_out/r2_bench_n${N}_d.json | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); e=d['e2e']; print('value', int(d['value']), 'ms', round(d['ms_per_step'],3), 'e2e', int(e['value']), round(e['ms_per_step'],1), e['learner_thread_ms_per_step'], e['actor_last_sample_ms'], e['copy_bandwidth_bidirectional_all_ranks_gbs'], int(e['pcie_ceiling_bidirectional_env_steps_per_s'] or 0))"
tail -3 gpurun_out/r2_bench_n${N}_d.err | cut -c1-300
