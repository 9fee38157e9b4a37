"""Where does the worst activation error of a model-level parity run sit?  (GPU; prints, asserts nothing.)
usage: python tools/diag_parity.py [model] [batch]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import vit_oracle as O  # noqa: E402
from vision_transformers_torch_xla_b200.models import create_model  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "vit_large_patch16_384"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 2
torch.backends.cuda.matmul.allow_tf32 = False
torch.backends.cudnn.allow_tf32 = False
dev = torch.device("cuda:0")
torch.manual_seed(0)
kw = dict(num_classes=1000, drop_path_rate=0.0)
if not name.startswith("deit_"):
    kw["global_pool"] = "avg"
ref = O.create_model(name, **kw).to(dev)
mine = create_model(name, **kw).to(dev)
mine.load_state_dict(ref.state_dict())
ref.eval()
mine.eval()
img = ref.patch_embed.img_size[0]
g = torch.Generator().manual_seed(3)
x = torch.randn(B, 3, img, img, generator=g).to(dev)
ins, outs = {"ref": [], "mine": []}, {"ref": [], "mine": []}
for tag, m in (("ref", ref), ("mine", mine)):
    m.blocks[0].register_forward_pre_hook(lambda mod, inp, t=tag: ins[t].append(inp[0].detach()))
    for blk in m.blocks:
        blk.register_forward_hook(lambda mod, inp, out, t=tag: outs[t].append(out.detach()))
with torch.no_grad():
    ref(x)
    mine(x)


def stats(tag, a, b):
    d = (a.double() - b.double())
    rms = b.double().pow(2).mean().sqrt()
    idx = d.abs().flatten().argmax().item()
    bb, n, c = idx // (a.shape[1] * a.shape[2]), (idx // a.shape[2]) % a.shape[1], idx % a.shape[2]
    row_rms = d.pow(2).mean(-1).sqrt() / rms            # [B, N]
    col_rms = d.pow(2).mean((0, 1)).sqrt() / rms        # [D]
    q = torch.quantile(d.abs().flatten()[:4_000_000].float() / float(rms), torch.tensor([0.5, 0.99, 0.9999], device=a.device))
    print(f"{tag}: rms(b) {float(rms):.4f}  rms err {float(d.pow(2).mean().sqrt() / rms):.3e}  max/rms {float(d.abs().max() / rms):.3e} at "
          f"(b={bb}, n={n}, d={c}) mine {float(a[bb, n, c]):+.5f} ref {float(b[bb, n, c]):+.5f}  |ref|/rms {abs(float(b[bb, n, c])) / float(rms):.2f}")
    print(f"     |err|/rms quantiles 50% {q[0]:.2e} 99% {q[1]:.2e} 99.99% {q[2]:.2e};  worst token row rms {float(row_rms.max()):.3e} at "
          f"{divmod(int(row_rms.flatten().argmax()), a.shape[1])} (median row {float(row_rms.median()):.3e});  worst channel rms "
          f"{float(col_rms.max()):.3e} at d={int(col_rms.argmax())} (median {float(col_rms.median()):.3e})")


stats("embed (input of block 0)", ins["mine"][0], ins["ref"][0])
for i in (0, 1, len(outs["ref"]) // 2, len(outs["ref"]) - 1):
    stats(f"block {i} output", outs["mine"][i], outs["ref"][i])
