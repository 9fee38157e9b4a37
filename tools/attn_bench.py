"""Times the attention forward / backward kernels at ViT-B shapes through the C ABI.  usage: attn_bench.py [B] [N] [H]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vision_transformers_torch_xla_b200 import _lib as L  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
N = int(sys.argv[2]) if len(sys.argv) > 2 else 197
H = int(sys.argv[3]) if len(sys.argv) > 3 else 12
ITERS = int(os.environ.get("GB_ITERS", "10"))
dev = torch.device("cuda")
torch.manual_seed(0)
qkv = torch.randn(B, N, 3 * H * 64, device=dev).bfloat16()
dout = torch.randn(B, N, H * 64, device=dev).bfloat16()
out = torch.empty(B, N, H * 64, device=dev, dtype=torch.bfloat16)
lse = torch.empty(B, H, N, device=dev)
dqkv = torch.empty_like(qkv)


def bench(name, fn, flops):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(ITERS):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / ITERS
    print(f"{name:28s} {ms * 1e3:8.1f} us  {flops / ms / 1e9:7.1f} TFLOP/s (algorithmic)", flush=True)


f = 4.0 * B * H * N * N * 64
print(f"B={B} N={N} H={H}  VITK_ATTN_FWD={os.environ.get('VITK_ATTN_FWD', 'persistent')} VITK_ATTN_BWD={os.environ.get('VITK_ATTN_BWD', '2wg')}")
bench("attn_fwd", lambda: L.attn_fwd(qkv, out, lse, B, N, H, 64, 0.125), f)
bench("attn_bwd", lambda: L.attn_bwd(qkv, out, dout, lse, dqkv, B, N, H, 64, 0.125), 2.5 * f)
q, k, v = qkv.view(B, N, 3, H, 64).permute(2, 0, 3, 1, 4).unbind(0)
bench("[torch SDPA fwd, library]", lambda: torch.nn.functional.scaled_dot_product_attention(q, k, v), f)
