"""PatchEmbed forward + weight gradient at the BASELINE shapes: the im2col-free path (bf16 NCHW image as a TMA operand)
against the explicit im2col path (vitk_patchify + GEMM on the materialised [B*P, 768] matrix).  CUDA events, 20 launches.
usage: python tools/patch_embed_bench.py [B] [img] [D]      env PE_ONLY=tma|patchify (for ncu captures)"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vision_transformers_torch_xla_b200 import _lib as L  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
IMG = int(sys.argv[2]) if len(sys.argv) > 2 else 224
D = int(sys.argv[3]) if len(sys.argv) > 3 else 768
ONLY = os.environ.get("PE_ONLY", "")
ITERS = int(os.environ.get("GB_ITERS", "20"))
dev = torch.device("cuda")
torch.manual_seed(0)
C, ps = 3, 16
gh = gw = IMG // ps
gwp = (gw + 7) // 8 * 8
P, K, prefix = gh * gw, C * ps * ps, 1
N = P + prefix
R = 3
imgs = [torch.randn(B, C, IMG, IMG, device=dev) for _ in range(R)]
imgb = [torch.empty(B, C, IMG, IMG, device=dev, dtype=torch.bfloat16) for _ in range(R)]
patches = [torch.empty(B * P, K, device=dev, dtype=torch.bfloat16) for _ in range(R)]
w = (torch.randn(D, K, device=dev) * 0.03).bfloat16()
bias, pos = torch.randn(D, device=dev), torch.randn(N, D, device=dev)
x = torch.empty(B, N, D, device=dev)
g = torch.randn(B, N, D, device=dev)
gp = torch.empty(B * P, D, device=dev, dtype=torch.bfloat16)
gpp = torch.empty(B * gh * gwp, D, device=dev, dtype=torch.bfloat16)
dW, db, dpos = torch.zeros(D, K, device=dev), torch.zeros(D, device=dev), torch.zeros(N, D, device=dev)
geom = (C, IMG, IMG, ps, gwp)


def timed(fn):
    for _ in range(3):
        fn(0)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(ITERS):
        fn(i)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / ITERS * 1e3


def tma_fwd(i):
    L.cast_bf16(imgs[i % R], imgb[i % R])
    L.gemm(imgb[i % R], w, x, M=B * gh * gwp, N=D, K=K, epilogue=L.EPI_PATCH, bias=bias, pos=pos, tokens_per_img=P, prefix=prefix,
           image=("a",) + geom)


def tma_fwd_bf16_input(i):
    L.gemm(imgb[i % R], w, x, M=B * gh * gwp, N=D, K=K, epilogue=L.EPI_PATCH, bias=bias, pos=pos, tokens_per_img=P, prefix=prefix,
           image=("a",) + geom)


def tma_bwd(i):
    L.embed_bwd(g, gpp, dpos, None, None, B, N, D, prefix, gw, gwp)
    L.gemm(gpp, imgb[i % R], dW, M=D, N=K, K=B * gh * gwp, epilogue=L.EPI_ATOMIC, a_mn=True, b_mn=True, colsum=db, image=("b",) + geom)


def im2col_fwd(i):
    L.patchify(imgs[i % R], patches[i % R], ps)
    L.gemm(patches[i % R], w, x, M=B * P, N=D, K=K, epilogue=L.EPI_PATCH, bias=bias, pos=pos, tokens_per_img=P, prefix=prefix)


def im2col_bwd(i):
    L.embed_bwd(g, gp, dpos, None, None, B, N, D, prefix)
    L.gemm(gp, patches[i % R], dW, M=D, N=K, K=B * P, epilogue=L.EPI_ATOMIC, a_mn=True, b_mn=True, colsum=db)


print(f"PatchEmbed B={B} {IMG}x{IMG} D={D}: P={P} patches/image, K={K}; padded patch-grid width {gwp} ({gw} real)")
if ONLY in ("", "tma"):
    print(f"  im2col-free  forward (fp32 image: cast + TMA GEMM)   {timed(tma_fwd):8.1f} us")
    print(f"  im2col-free  forward (bf16 image: TMA GEMM only)     {timed(tma_fwd_bf16_input):8.1f} us")
    print(f"  im2col-free  backward (embed_bwd + TMA wgrad)        {timed(tma_bwd):8.1f} us")
if ONLY in ("", "patchify"):
    print(f"  explicit im2col forward (patchify + GEMM)            {timed(im2col_fwd):8.1f} us")
    print(f"  explicit im2col backward (embed_bwd + wgrad)         {timed(im2col_bwd):8.1f} us")
