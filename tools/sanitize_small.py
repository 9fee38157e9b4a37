"""Small-shape exercise of the kernels added in round 2 (written for `compute-sanitizer --tool memcheck`; the tool is closed on
this GPU pool, so it serves as a quick stand-alone smoke of the new paths):
attn_fwd6 (N = 300), the wide-head forward / streaming backward (head_dim 72), the streaming backward (N = 300), the GELU /
RESID epilogues with a dropout mask, the image-operand PatchEmbed GEMM, dropout_mask / mask_mul, droppath masks."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vision_transformers_torch_xla_b200 import _lib as L  # noqa: E402

dev = torch.device("cuda")
torch.manual_seed(0)


def attn(B, N, H, hd):
    qkv = torch.randn(B, N, 3 * H * hd, device=dev).bfloat16()
    out = torch.empty(B, N, H * hd, device=dev, dtype=torch.bfloat16)
    lse = torch.empty(B, H, N, device=dev)
    dout = torch.randn(B, N, H * hd, device=dev).bfloat16()
    dqkv = torch.empty_like(qkv)
    L.attn_fwd(qkv, out, lse, B, N, H, hd, hd ** -0.5)
    L.attn_bwd(qkv, out, dout, lse, dqkv, B, N, H, hd, hd ** -0.5)
    torch.cuda.synchronize()
    assert torch.isfinite(out.float()).all() and torch.isfinite(dqkv.float()).all()
    print(f"attention B={B} N={N} H={H} hd={hd} ok", flush=True)


attn(2, 300, 2, 64)     # attn_fwd6 + streaming backward
attn(1, 577, 1, 48)     # narrow heads, odd number of q tiles (idle group in the last round)
attn(2, 197, 2, 72)     # wide heads: [tile][tail] operands
attn(2, 197, 2, 64)     # attn_fwd4 / attn_bwd4
attn(1, 100, 1, 80)

M, N, K = 300, 384, 128
x = torch.randn(M, K, device=dev).bfloat16()
w = torch.randn(N, K, device=dev).bfloat16()
b = torch.randn(N, device=dev)
m = (torch.rand(M, N, device=dev) > 0.2).to(torch.uint8)
out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
aux = torch.empty_like(out)
L.gemm(x, w, out, M=M, N=N, K=K, epilogue=L.EPI_GELU, bias=b, aux=aux, mask=m, mask_scale=1.25)
r = torch.randn(M, N, device=dev)
o32 = torch.empty(M, N, device=dev)
L.gemm(x, w, o32, M=M, N=N, K=K, epilogue=L.EPI_RESID, bias=b, resid=r, mask=m, mask_scale=1.25)
torch.cuda.synchronize()
print("gemm epilogues with a dropout mask ok", flush=True)

mask = torch.empty(1003, dtype=torch.uint8, device=dev)
L.dropout_mask(mask, 0.3, 1, 2)
y = torch.randn(40, 64, device=dev).bfloat16()
L.mask_mul_(y, (torch.rand(40, 64, device=dev) > 0.5).to(torch.uint8), 2.0, 40, 64)
z = torch.randn(40, 64, device=dev)
L.mask_mul_(z, (torch.rand(40, 64, device=dev) > 0.5).to(torch.uint8), 2.0, 40, 64)
rs = torch.empty(4, 6, device=dev)
L.droppath_masks(rs, [0.0, 0.1, 0.2, 0.3], 5, 1)
torch.cuda.synchronize()
print("dropout / droppath kernels ok", flush=True)

from vision_transformers_torch_xla_b200.losses import SoftTargetCrossEntropy  # noqa: E402
from vision_transformers_torch_xla_b200.models import create_model  # noqa: E402

model = create_model("my_vit_xs", num_classes=16, global_pool="avg", drop_path_rate=0.1, proj_drop_rate=0.1).to(dev).train()
img = torch.randn(2, 3, 224, 224, device=dev)
tgt = torch.softmax(torch.randn(2, 16, device=dev), -1)
SoftTargetCrossEntropy()(model(img), tgt).backward()
torch.cuda.synchronize()
print("my_vit_xs step (image-operand PatchEmbed, wide heads, dropout, DropPath) ok", flush=True)
