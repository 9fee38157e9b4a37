"""Times the ViT-B GEMM shapes of one training step through the C ABI (CUDA events, L2-cold rotation of buffers).
usage: python tools/gemm_bench.py [M]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vision_transformers_torch_xla_b200 import _lib as L  # noqa: E402

M = int(sys.argv[1]) if len(sys.argv) > 1 else 50432
ONLY = sys.argv[2] if len(sys.argv) > 2 else ""   # substring filter, e.g. "fc1 fprop"
ITERS = int(os.environ.get("GB_ITERS", "20"))
D, F = 768, 3072
dev = torch.device("cuda")
torch.manual_seed(0)


def bench(name, fn, flops, iters=None):
    iters = iters or ITERS
    if ONLY and ONLY not in name:
        return
    for _ in range(3):
        fn(0)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(iters):
        fn(i)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    print(f"{name:34s} {ms * 1e3:8.1f} us  {flops / ms / 1e9:8.1f} TFLOP/s", flush=True)


R = 3  # rotate buffers so operands do not sit in L2 between iterations
xD = [torch.randn(M, D, device=dev).bfloat16() for _ in range(R)]
xF = [torch.randn(M, F, device=dev).bfloat16() for _ in range(R)]
x3 = [torch.randn(M, 3 * D, device=dev).bfloat16() for _ in range(R)]
oF = [torch.empty(M, F, device=dev, dtype=torch.bfloat16) for _ in range(R)]
aF = [torch.empty(M, F, device=dev, dtype=torch.bfloat16) for _ in range(R)]
o3 = [torch.empty(M, 3 * D, device=dev, dtype=torch.bfloat16) for _ in range(R)]
oD = [torch.empty(M, D, device=dev, dtype=torch.bfloat16) for _ in range(R)]
rD = [torch.randn(M, D, device=dev) for _ in range(R)]
fD = [torch.empty(M, D, device=dev) for _ in range(R)]
Wqkv = (torch.randn(3 * D, D, device=dev) * 0.02).bfloat16()
W1 = (torch.randn(F, D, device=dev) * 0.02).bfloat16()
W2 = (torch.randn(D, F, device=dev) * 0.02).bfloat16()
Wp = (torch.randn(D, D, device=dev) * 0.02).bfloat16()
b3, bF, bD = torch.randn(3 * D, device=dev), torch.randn(F, device=dev), torch.randn(D, device=dev)
gW1, gW2 = torch.zeros(F, D, device=dev), torch.zeros(D, F, device=dev)
gWq, gWp = torch.zeros(3 * D, D, device=dev), torch.zeros(D, D, device=dev)

print(f"M={M}  VITK_GEMM_ROLES={os.environ.get('VITK_GEMM_ROLES', 'hi')}")
bench("qkv fprop  bias->bf16", lambda i: L.gemm(xD[i % R], Wqkv, o3[i % R], M=M, N=3 * D, K=D, epilogue=L.EPI_BF16, bias=b3), 2 * M * 3 * D * D)
bench("fc1 fprop  bias+GELU (2 outs)", lambda i: L.gemm(xD[i % R], W1, oF[i % R], M=M, N=F, K=D, epilogue=L.EPI_GELU, bias=bF, aux=aF[i % R]), 2 * M * F * D)
bench("fc2 fprop  bias+resid f32", lambda i: L.gemm(xF[i % R], W2, fD[i % R], M=M, N=D, K=F, epilogue=L.EPI_RESID, bias=bD, resid=rD[i % R]), 2 * M * F * D)
bench("proj fprop bias+resid f32", lambda i: L.gemm(xD[i % R], Wp, fD[i % R], M=M, N=D, K=D, epilogue=L.EPI_RESID, bias=bD, resid=rD[i % R]), 2 * M * D * D)
bench("fc2 dgrad  x gelu' -> bf16", lambda i: L.gemm(xD[i % R], W2, oF[i % R], M=M, N=F, K=D, epilogue=L.EPI_DGELU, b_mn=True, aux=aF[i % R]), 2 * M * F * D)
bench("fc1 dgrad  -> bf16", lambda i: L.gemm(xF[i % R], W1, oD[i % R], M=M, N=D, K=F, epilogue=L.EPI_BF16, b_mn=True), 2 * M * F * D)
bench("qkv dgrad  -> bf16", lambda i: L.gemm(x3[i % R], Wqkv, oD[i % R], M=M, N=D, K=3 * D, epilogue=L.EPI_BF16, b_mn=True), 2 * M * 3 * D * D)
bench("proj dgrad -> bf16", lambda i: L.gemm(xD[i % R], Wp, oD[i % R], M=M, N=D, K=D, epilogue=L.EPI_BF16, b_mn=True), 2 * M * D * D)
bench("fc1 wgrad  split-K red.add", lambda i: L.gemm(xF[i % R], xD[i % R], gW1, M=F, N=D, K=M, epilogue=L.EPI_ATOMIC, a_mn=True, b_mn=True), 2 * M * F * D)
bench("fc2 wgrad  split-K red.add", lambda i: L.gemm(xD[i % R], xF[i % R], gW2, M=D, N=F, K=M, epilogue=L.EPI_ATOMIC, a_mn=True, b_mn=True), 2 * M * F * D)
bench("qkv wgrad  split-K red.add", lambda i: L.gemm(x3[i % R], xD[i % R], gWq, M=3 * D, N=D, K=M, epilogue=L.EPI_ATOMIC, a_mn=True, b_mn=True), 2 * M * 3 * D * D)
bench("proj wgrad split-K red.add", lambda i: L.gemm(xD[i % R], xD[(i + 1) % R], gWp, M=D, N=D, K=M, epilogue=L.EPI_ATOMIC, a_mn=True, b_mn=True), 2 * M * D * D)
# library reference point (NOT used by the product): cuBLAS bf16 on the same shapes
bench("[cuBLAS] fc1 fprop plain", lambda i: torch.matmul(xD[i % R], W1.t(), out=oF[i % R]), 2 * M * F * D)
bench("[cuBLAS] fc2 fprop plain", lambda i: torch.matmul(xF[i % R], W2.t(), out=oD[i % R]), 2 * M * F * D)
bench("[cuBLAS] qkv fprop plain", lambda i: torch.matmul(xD[i % R], Wqkv.t(), out=o3[i % R]), 2 * M * 3 * D * D)
