"""Prints the phase timeline (SM cycles) of the first CTA of the attention kernels.  usage: attn_trace.py [B] [N] [H]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vision_transformers_torch_xla_b200 import _lib as L  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
N = int(sys.argv[2]) if len(sys.argv) > 2 else 197
H = int(sys.argv[3]) if len(sys.argv) > 3 else 12
dev = torch.device("cuda")
qkv = torch.randn(B, N, 3 * H * 64, device=dev).bfloat16()
dout = torch.randn(B, N, H * 64, device=dev).bfloat16()
out = torch.empty(B, N, H * 64, device=dev, dtype=torch.bfloat16)
lse = torch.empty(B, H, N, device=dev)
dqkv = torch.empty_like(qkv)
trace = torch.zeros(64, dtype=torch.int64, device=dev)
lib = L.load()
for _ in range(2):
    L.attn_fwd(qkv, out, lse, B, N, H, 64, 0.125)
    L.attn_bwd(qkv, out, dout, lse, dqkv, B, N, H, 64, 0.125)
for name, fn in (("attn_fwd", lambda: L.attn_fwd(qkv, out, lse, B, N, H, 64, 0.125)),
                 ("attn_bwd", lambda: L.attn_bwd(qkv, out, dout, lse, dqkv, B, N, H, 64, 0.125))):
    trace.zero_()
    lib.vitk_debug_set_trace(trace.data_ptr())
    fn()
    torch.cuda.synchronize()
    lib.vitk_debug_set_trace(None)
    t = trace.cpu().tolist()
    t0 = t[0]
    print(f"== {name}: first CTA, thread 0, cycles since kernel entry (slot: cycles, delta)")
    prev = t0
    for i, v in enumerate(t):
        if v:
            print(f"  slot {i:2d}: {v - t0:8d}  (+{v - prev})")
            prev = v
