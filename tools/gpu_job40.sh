#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout -s KILL 60 python -m pytest tests/test_gpu_vtrace.py -q -x 2>&1 | tail -2
timeout -s KILL 80 python tools/k1_promo.py 2>&1 | tee gpurun_out/r2_k1_promo.jsonl | tail -14
