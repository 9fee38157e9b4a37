"""Times the LayerNorm forward / backward kernels at ViT-B shapes through the C ABI and prints achieved HBM GB/s against
the algorithmic bytes (fwd 6*D, bwd 16*D bytes per row).  usage: ln_bench.py [rows] [dim]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vision_transformers_torch_xla_b200 import _lib as L  # noqa: E402

M = int(sys.argv[1]) if len(sys.argv) > 1 else 50432
D = int(sys.argv[2]) if len(sys.argv) > 2 else 768
ITERS = int(os.environ.get("GB_ITERS", "20"))
dev = torch.device("cuda")
R = 3  # rotate buffers: every iteration streams from HBM, not L2
x = [torch.randn(M, D, device=dev) for _ in range(R)]
dy = [torch.randn(M, D, device=dev).bfloat16() for _ in range(R)]
gin = [torch.randn(M, D, device=dev) for _ in range(R)]
gout = [torch.empty(M, D, device=dev) for _ in range(R)]
gb = [torch.empty(M, D, device=dev, dtype=torch.bfloat16) for _ in range(R)]
y = [torch.empty(M, D, device=dev, dtype=torch.bfloat16) for _ in range(R)]
w, b = torch.ones(D, device=dev), torch.zeros(D, device=dev)
dw, db = torch.zeros(D, device=dev), torch.zeros(D, device=dev)
mean, rstd = torch.empty(M, device=dev), torch.empty(M, device=dev)
rsc = torch.ones(M // 197 + 1, device=dev)


def bench(name, fn, nbytes):
    for i in range(3):
        fn(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(ITERS):
        fn(i)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / ITERS
    print(f"{name:28s} {ms * 1e3:8.1f} us  {nbytes / ms / 1e6:8.1f} GB/s (algorithmic)", flush=True)


bench("layernorm_fwd", lambda i: L.layernorm_fwd(x[i % R], w, b, y[i % R], mean, rstd, M, D, 1e-6), 6.0 * M * D)
bench("layernorm_bwd (+resid, +bf16)", lambda i: L.layernorm_bwd(dy[i % R], x[i % R], mean, rstd, w, gin[i % R], gout[i % R], gb[i % R], rsc, 197, dw, db, M, D), 16.0 * M * D)
c = torch.empty(M * D * 4, device=dev, dtype=torch.uint8)
c2 = torch.empty_like(c)
bench("[torch copy, library]", lambda i: c2.copy_(c), 2.0 * M * D * 4)
