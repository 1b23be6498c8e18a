#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout -s KILL 300 python -m pytest tests/test_gpu_gemm.py -x -q 2>&1 | tail -5
for rep in 1 2; do for h in 0 1; do
echo "rep $rep FUSE_HEADS=$h (mma head kernel)"; PARL_B200_FUSE_HEADS=$h timeout -s KILL 300 python tools/impala_phases.py 512 4096 2>&1 | tail -2
done; done
