"""GPU: the "library-kernel bar" of SURVEY section 8(d) — the same training step (BASELINE configs 3, 2 and 5: ViT-B/16 and
ViT-S/16 at batch 256, ViT-L/16 at 384 px and batch 64; drop_path 0.1, SoftTargetCE, AdamW) run by PyTorch's own kernels on the same B200: the fp32 oracle under ``torch.autocast(bfloat16)``
(cuBLAS GEMMs, fused SDPA, eager LayerNorm / GELU / residual adds) with ``torch.optim.AdamW(fused=True)``.  It is a reported
reference point next to the CPU baseline of bench.py, not a target; the test only requires the hand-written path to be the
faster of the two and prints both numbers (run with ``-s``; profiles/r02_library_bar.txt holds a recorded run).
Timed on the device with CUDA events, 3 warm-up + 10 timed steps each, same synthetic batch, inputs resident in HBM."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _time(step, warmup=3, steps=10):
    for _ in range(warmup):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        step()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


@pytest.mark.parametrize("name,B,img", [("vit_base_patch16_224", 256, 224), ("vit_small_patch16_224", 256, 224),
                                        ("vit_large_patch16_384", 64, 384)])
def test_library_bar(cuda_device, name, B, img):
    from oracle import vit_oracle as O
    from vision_transformers_torch_xla_b200 import optim_factory
    from vision_transformers_torch_xla_b200.losses import SoftTargetCrossEntropy
    from vision_transformers_torch_xla_b200.models import create_model

    dev = cuda_device
    kw = dict(num_classes=1000, global_pool="avg", drop_path_rate=0.1)
    g = torch.Generator().manual_seed(0)
    x = torch.randn(B, 3, img, img, generator=g).to(dev)
    tgt = O.mixup_soft_targets(torch.randint(0, 1000, (B,), generator=g)).to(dev)

    torch.manual_seed(0)
    ref = O.create_model(name, **kw).to(dev).train()
    opt_ref = torch.optim.AdamW(ref.parameters(), lr=1e-3, weight_decay=0.05, fused=True)
    crit_ref = O.SoftTargetCrossEntropy()

    def lib_step():
        with torch.autocast("cuda", dtype=torch.bfloat16):
            out = ref(x)
        crit_ref(out.float(), tgt).backward()
        opt_ref.step()
        opt_ref.zero_grad(set_to_none=True)

    lib_ms = _time(lib_step)
    del ref, opt_ref
    torch.cuda.empty_cache()

    class Args:
        opt, lr, weight_decay, opt_eps, opt_betas = "adamw", 1e-3, 0.05, 1e-8, None

    torch.manual_seed(0)
    mine = create_model(name, **kw).to(dev).train()
    opt = optim_factory.create_optimizer(Args, mine)
    crit = SoftTargetCrossEntropy()

    def vitk_step():
        crit(mine(x), tgt).backward()
        opt.step()
        opt.zero_grad()

    vitk_ms = _time(vitk_step)
    print(f"\n[bar] {name} batch {B}, one B200: torch eager autocast(bf16) + fused AdamW {lib_ms:.2f} ms/step = {B / lib_ms * 1e3:.0f} img/s"
          f" | vitk {vitk_ms:.2f} ms/step = {B / vitk_ms * 1e3:.0f} img/s | {lib_ms / vitk_ms:.2f}x")
    assert vitk_ms < lib_ms
