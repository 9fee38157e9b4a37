"""Timeline (SM cycles) of the second work item of CTA 0 of the long-sequence attention forward (attn_fwd5).
usage: attn_trace5.py [B] [N] [H]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vision_transformers_torch_xla_b200 import _lib as L  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
N = int(sys.argv[2]) if len(sys.argv) > 2 else 577
H = int(sys.argv[3]) if len(sys.argv) > 3 else 16
dev = torch.device("cuda")
qkv = torch.randn(B, N, 3 * H * 64, device=dev).bfloat16()
out = torch.empty(B, N, H * 64, device=dev, dtype=torch.bfloat16)
lse = torch.empty(B, H, N, device=dev)
trace = torch.zeros(128, dtype=torch.int64, device=dev)
lib = L.load()
for _ in range(2):
    L.attn_fwd(qkv, out, lse, B, N, H, 64, 0.125)
lib.vitk_debug_set_trace(trace.data_ptr())
L.attn_fwd(qkv, out, lse, B, N, H, 64, 0.125)
torch.cuda.synchronize()
lib.vitk_debug_set_trace(None)
t = trace.cpu().tolist()
ev = []
for g in (0, 1):
    for i in range(16):
        if t[16 * g + i]:
            ev.append((t[16 * g + i], f"softmax g{g}: step {i // 2} " + ("S ready" if i % 2 == 0 else "P written (O rescaled)")))
        if t[32 + 16 * g + i]:
            ev.append((t[32 + 16 * g + i], f"    mma g{g}: step {i // 2} " + ("S issued" if i % 2 == 0 else "P ready")))
for j in range(3):
    for k, nm in enumerate(['S in registers', 'max done', 'exp + P stores issued', 'O rescaled, stores landed']):
        if t[64 + j * 4 + k]:
            ev.append((t[64 + j * 4 + k], f'softmax g0: step {j}    {nm}'))
ev.sort()
t0 = ev[0][0] if ev else 0
for v, name in ev:
    print(f"{v - t0:8d}  {name}")
