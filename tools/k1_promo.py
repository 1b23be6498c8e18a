"""K1 v8: L2 promotion of the logits tensor maps (128 B default / 256 B / none), graph-timed like bench_k1."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from parl_b200 import _lib  # noqa
from tools.bench_kernels import bench_vtrace  # noqa

lib = _lib.load()
for (T, B, A) in [(50, 4096, 18), (50, 65536, 18)]:
    for rep in range(2):
        for mode in (0, 10, 11):
            lib.rl_debug_set_vtrace_path(mode)
            r = bench_vtrace(T, B, A, nbuf=8 if B <= 4096 else 2)
            print(json.dumps(dict(B=B, mode=mode, promo={0: '128B', 10: '256B', 11: 'none'}[mode], us=round(r['us'], 3),
                                  frac=round(r['gbps'] / 6571.6, 4))))
lib.rl_debug_set_vtrace_path(0)
