"""Per-kernel timings through the C ABI (CUDA events, L2 flushed between iterations).  Not a pytest file.
    python tests/microbench.py > gpurun_out/microbench.log
"""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cgvc  # noqa: E402
from cgvc import native as N  # noqa: E402

lib = N.load()
cfg = N.Config(24, 1, 128, 0, 0, 0)
h = C.c_void_p(0)
assert lib.cgvc_create(C.byref(cfg), C.byref(h)) == 0
P = lambda t: C.c_void_p(t.data_ptr())
flush = torch.empty(64 * 1024 * 1024, dtype=torch.float32, device="cuda")   # 256 MB > L2


def timeit(fn, iters=5):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return sorted(ts)[len(ts) // 2]


print("== IN+GLU post kernels (fp32 in/out through cgvc_in_glu_*), batch 512")
for (B, R, Cc, sh) in [(512, 32, 1024, 1), (512, 64, 512, 2), (512, 128, 256, 2), (512, 384, 256, 1), (512, 96, 512, 1), (512, 48, 1024, 1), (512, 64, 256, 1)]:
    p = torch.randn(B, R // sh, 2 * Cc * sh, device="cuda"); prm = [torch.randn(Cc, device="cuda") for _ in range(4)]
    y = torch.empty(B, R, Cc, device="cuda"); st = torch.empty(B, 4, Cc, device="cuda"); dy = torch.randn(B, R, Cc, device="cuda")
    dp = torch.empty_like(p); g = [torch.zeros(Cc, device="cuda") for _ in range(4)]
    tf = timeit(lambda: lib.cgvc_in_glu_forward(h, P(p), P(prm[0]), P(prm[1]), P(prm[2]), P(prm[3]), P(y), P(st), B, R, Cc, sh, None))
    tb = timeit(lambda: lib.cgvc_in_glu_backward(h, P(dy), P(p), P(st), P(prm[0]), P(prm[1]), P(prm[2]), P(prm[3]), P(dp), P(g[0]), P(g[1]), P(g[2]), P(g[3]), B, R, Cc, sh, None))
    ne = B * R * Cc
    print("B=%d R=%d C=%d sh=%d  fwd %.3f ms (%.0f GB/s of 12 B/elem)  bwd %.3f ms (%.0f GB/s of 20 B/elem)"
          % (B, R, Cc, sh, tf, ne * 12 / tf / 1e6, tb, ne * 20 / tb / 1e6))

print("== tensor-core convs through cgvc_conv_forward/backward (includes temporary plane building: upper bound only)")
lib.cgvc_destroy(h)
