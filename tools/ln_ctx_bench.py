"""LayerNorm backward timed in three contexts: back to back with itself, right after a large bf16 GEMM (the in-step
situation: power-capped clocks, the producer's dirty lines in L2), and after a GEMM plus an L2-sized memset.
usage: ln_ctx_bench.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vision_transformers_torch_xla_b200 import _lib as L  # noqa: E402

M, D = 50432, 768
dev = torch.device("cuda")
R = 3
x = [torch.randn(M, D, device=dev) for _ in range(R)]
dy = [torch.randn(M, D, device=dev).bfloat16() for _ in range(R)]
g = [torch.randn(M, D, device=dev) for _ in range(R)]
gb = [torch.empty(M, D, device=dev, dtype=torch.bfloat16) for _ in range(R)]
w = torch.ones(D, device=dev)
dw, db = torch.zeros(D, device=dev), torch.zeros(D, device=dev)
mean, rstd = torch.zeros(M, device=dev), torch.ones(M, device=dev)
rsc = torch.ones(M // 197 + 1, device=dev)
a = torch.randn(8192, 8192, device=dev).bfloat16()
b = torch.randn(8192, 8192, device=dev).bfloat16()
c = torch.empty(8192, 8192, device=dev, dtype=torch.bfloat16)
flush = torch.empty(256 << 20, device=dev, dtype=torch.uint8)
src = torch.empty(M * D * 4, device=dev, dtype=torch.uint8)
dst = torch.empty_like(src)


def ln(i):
    L.layernorm_bwd(dy[i % R], x[i % R], mean, rstd, w, g[i % R], g[i % R], gb[i % R], rsc, 197, dw, db, M, D)


y = [torch.empty(M, D, device=dev, dtype=torch.bfloat16) for _ in range(R)]
bb = torch.zeros(D, device=dev)


def lnf(i):
    L.layernorm_fwd(x[i % R], w, bb, y[i % R], mean, rstd, M, D, 1e-6)


def cp(i):
    dst.copy_(src)


def run(name, fn, nbytes, pre, iters=30):
    evs = []
    for i in range(iters + 5):
        pre()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn(i)
        e1.record()
        evs.append((e0, e1))
    torch.cuda.synchronize()
    ts = sorted(e0.elapsed_time(e1) for e0, e1 in evs[5:])
    med = ts[len(ts) // 2]
    print(f"{name:50s} median {med * 1e3:7.1f} us  {nbytes / med / 1e6:7.0f} GB/s", flush=True)


def gemms(n):
    def f():
        for _ in range(n):
            torch.matmul(a, b, out=c)
    return f


def gemms_flush(n):
    def f():
        for _ in range(n):
            torch.matmul(a, b, out=c)
        flush.zero_()
    return f


for name, fn, nb in (("layernorm_bwd", ln, 16.0 * M * D), ("layernorm_fwd", lnf, 6.0 * M * D), ("torch copy 155 MB", cp, 8.0 * M * D)):
    run(name + " | back to back", fn, nb, lambda: None)
    run(name + " | after 1 GEMM 8192^3", fn, nb, gemms(1))
    run(name + " | after 20 GEMMs (power-capped clocks)", fn, nb, gemms(20), iters=20)
    run(name + " | after 20 GEMMs + 256 MB memset", fn, nb, gemms_flush(20), iters=20)
