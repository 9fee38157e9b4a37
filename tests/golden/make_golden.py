"""Generates tests/golden/vit_micro.pt from the CPU oracle (seeded, deterministic).

The reference itself cannot be imported offline (timm/torch_xla absent), so these vectors pin the
ORACLE (regression pin) and give the GPU kernels a committed fixture to match; they are not outputs of the
reference.  Run from the repo root:  python tests/golden/make_golden.py
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import vit_oracle as O  # noqa: E402

CFG = dict(img_size=32, patch_size=16, embed_dim=64, depth=2, num_heads=1, num_classes=16)


def make(global_pool):
    torch.manual_seed(1234)
    torch.set_num_threads(1)
    model = O.VisionTransformer(global_pool=global_pool, **CFG)
    # non-trivial LN / bias values so that every parameter matters
    with torch.no_grad():
        for n, p in model.named_parameters():
            if n.endswith("bias") or "norm" in n:
                p.add_(0.1 * torch.randn_like(p))
    model.train()
    x = torch.randn(3, 3, 32, 32)
    tgt = torch.softmax(torch.randn(3, 16) * 2, -1)
    logits = model(x)
    loss = O.SoftTargetCrossEntropy()(logits, tgt)
    loss.backward()
    grads = {n: p.grad.clone() for n, p in model.named_parameters()
             if n in ("cls_token", "pos_embed", "patch_embed.proj.weight", "blocks.0.attn.qkv.weight",
                      "blocks.0.norm1.weight", "blocks.1.mlp.fc2.bias", "head.weight",
                      "fc_norm.weight" if global_pool == "avg" else "norm.weight")}
    return dict(cfg=dict(CFG, global_pool=global_pool), state_dict={k: v.clone() for k, v in model.state_dict().items()},
                x=x, target=tgt, logits=logits.detach(), loss=loss.detach(), grads=grads)


if __name__ == "__main__":
    out = {gp: make(gp) for gp in ("avg", "token")}
    # both pools share every weight except the name of the final norm: store the state_dict once
    sd = out["avg"].pop("state_dict")
    tok = out["token"].pop("state_dict")
    assert all(torch.equal(sd[k.replace("norm.", "fc_norm.") if k.startswith("norm.") else k], v) for k, v in tok.items())
    out["state_dict_avg"] = sd
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "vit_micro.pt")
    torch.save(out, path)
    print("wrote", path, os.path.getsize(path), "bytes")
