#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout -s KILL 500 python -m pytest tests/test_gpu_exact.py tests/test_gpu_gemm.py tests/test_gpu_conv1_u8.py tests/test_gpu_trainnet.py -x -q 2>&1 | tail -5
timeout -s KILL 200 python tools/conv_forms.py > gpurun_out/r2_conv_forms_b.jsonl 2> gpurun_out/r2_conv_forms_b.err; cat gpurun_out/r2_conv_forms_b.jsonl; tail -3 gpurun_out/r2_conv_forms_b.err
for rep in 1 2; do
for dt in bf16 uint8; do
PARL_B200_OBS_DTYPE=$dt timeout -s KILL 300 python bench.py --steps 15 --warmup 4 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('obs $dt rep $rep ms_per_step',round(d['ms_per_step'],3),'value',int(d['value']),'conv1 frac',round(d['roofline']['frac'],3),'us',round(d['roofline']['us_per_launch'],1))"
done; done
