"""rl_gemm_bf16_tn: single-CTA form vs 2 x 2 cluster + TMA multicast form vs torch.matmul (cuBLAS) at the fc shapes of
the Atari network (actor batch and learner batch); checks that the two forms agree bit for bit."""
import json
import sys

import torch

sys.path.insert(0, __file__.rsplit('/', 2)[0])
from parl_b200 import kernels as K  # noqa: E402

dev = torch.device('cuda', 0)
bf = torch.bfloat16
torch.manual_seed(0)


def timeit(fn, reps=8):
    ev = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        ev.append((a, b))
    torch.cuda.synchronize()
    ms = sorted(x.elapsed_time(y) for x, y in ev[1:])
    return ms[len(ms) // 2] * 1e3


shapes = [(4096, 512, 5184), (1024, 512, 5184), (204800, 512, 5184), (204800, 5184, 512)]
if '--small' in sys.argv:
    shapes = shapes[:2]
for M, N, Kd in shapes:
    a = (torch.randn(M, Kd, device=dev) * 0.1).to(bf)
    b = (torch.randn(N, Kd, device=dev) * 0.1).to(bf)
    bias = torch.randn(N, device=dev)
    out = {}
    res = dict(M=M, N=N, K=Kd)
    for cl in (0, 1):
        K.set_gemm_cluster(cl)
        o = torch.empty(M, N, device=dev, dtype=bf)
        fn = lambda: K.gemm_bf16_tn(a, b, bias, relu=True, out=o)
        fn()
        torch.cuda.synchronize()
        out[cl] = o.clone()
        res['cluster_us' if cl else 'single_us'] = timeit(fn)
    K.set_gemm_cluster(1)
    res['forms_equal'] = bool(torch.equal(out[0], out[1]))
    bt = b.t().contiguous()
    o2 = torch.empty(M, N, device=dev, dtype=bf)
    res['cublas_us'] = timeit(lambda: torch.matmul(a, bt, out=o2))
    ref = torch.relu(o2.float() + bias)
    res['max_abs_diff_vs_cublas'] = float((out[1].float() - ref).abs().max())
    fl = 2.0 * M * N * Kd
    for k in ('single_us', 'cluster_us', 'cublas_us'):
        res[k.replace('_us', '_tflops')] = fl / (res[k] * 1e-6) / 1e12
    print(json.dumps(res))
    del a, b, out, o2, bt, ref
    torch.cuda.empty_cache()
