#!/bin/bash
# full GPU suite + smoke
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
echo "== pytest cc first"; timeout -s KILL 300 python -m pytest tests/test_gpu_cc.py -q 2>&1 | tail -15
echo "== pytest -m gpu (all)"; timeout -s KILL 1200 python -m pytest tests -m gpu -q --durations=8 2>&1 | tail -25
echo "== smoke"; timeout -s KILL 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
