"""GPU: whole-model parity on the BASELINE configs themselves (not only ViT-Ti): ViT-S/16, ViT-B/16 (the headline),
DeiT-B distilled (N = 198, distilled_training on), ViT-L/16 at 384 px (D = 1024, N = 577, 24 blocks: the kv-loop
attention forward, the streaming attention backward and the 8-warp LayerNorm ring), all with drop_path_rate = 0.1 — the
value the reference launches with (run_train.sh:58) and the one bench.py times.

Reference side: the fp32 oracle on the same GPU (TF32 off), same weights, same batch.  DropPath: the oracle draws its
masks with ``bernoulli_`` as timm does; they are recorded and REPLAYED into the CUDA path (ops.mask_source), so the
rs1 / rs2 / prev_rs hand-off between the fused stages is compared exactly.

Stated tolerances (north_star: bf16 vs fp32 reference, max rel err <= 2e-2).  Three metrics per tensor, all printed:
  elem  = max |a-b| / (|b| + rms(b))   the relative error of the worst element, with an RMS floor for near-zero elements
  rms   = rms(a-b) / rms(b)            the typical error
  max/rms = max |a-b| / rms(b)         the worst ABSOLUTE error in units of the tensor's RMS
  first block's output                elem <= 2e-2
  every block's output, logits       elem <= 3e-2, rms <= 7e-3, max/rms <= 4e-2
                                     AND elem <= 1.25 x the error of the SAME oracle run under torch.autocast(bfloat16)
                                     (PyTorch's own bf16 path: cuBLAS GEMMs + SDPA, fp32 residual adds) + 2e-3
  residual-stream gradients          rms <= 1e-2 and cosine >= 0.9999; worst-element metrics printed, and bounded (elem <= 8e-2,
                                     max/rms <= 1e-1) for the 'avg'-pooled models only: with 'token' pooling the loss reaches
                                     the last blocks' stream through two token rows per image, so the tensor's RMS says
                                     nothing about its large elements (DeiT-B: rms 7.9e-3, max/rms 0.35)
  every parameter gradient           rms <= 1.5e-2 and cosine >= 0.999  (2e-2 for the 24-block ViT-L), max/rms printed
Why the worst element of a deep model's activations is not held to 2e-2 (measured, tools/diag_parity.py,
profiles/r02_parity_diag.txt): the error is homogeneous bf16 operand-rounding noise - 2.4e-3 of the RMS after the patch
embedding alone (one K = 768 GEMM), 3.0e-3 after block 0, 5.0e-3 after 12 blocks, 5.9e-3 after 24, no token row or
channel standing out - and a [8, 197, 768] tensor has 1.2 M elements whose worst one sits 5-6 sigma out: the 99.99 %
quantile of |err| / rms is 2.0e-2 at ViT-B's last block, the maximum 2.7e-2.  Only higher-precision GEMM operands would
move that; the yardstick that says the kernels add nothing avoidable is the library's own bf16 run, asserted above.
ViT-Ti (300 K-element tensors, 12 blocks of D = 192) stays below 2e-2 in all three metrics (tests/test_gpu_model.py).
"""
import pytest
import torch

from conftest import cos_sim, elem_err, rel_err, report, rms_err

pytestmark = pytest.mark.gpu


class record_masks:
    """Collects the DropPath masks the oracle draws (as mask / keep_prob factors, in draw order)."""

    def __init__(self, model):
        from oracle import vit_oracle as O

        self.masks, self.hooks = [], []
        for mod in model.modules():
            if isinstance(mod, O.DropPath) and mod.drop_prob > 0:
                self.hooks.append(mod.register_forward_hook(self._hook))

    def _hook(self, mod, inp, out):
        x = inp[0]
        flat_x, flat_o = x.reshape(x.shape[0], -1), out.reshape(x.shape[0], -1)
        j = flat_x.abs().argmax(dim=1, keepdim=True)          # a safely non-zero element of every sample
        self.masks.append((flat_o.gather(1, j) / flat_x.gather(1, j)).flatten().detach().float())

    def close(self):
        for h in self.hooks:
            h.remove()


def replay(masks):
    def source(drop_probs, B, device):
        it = iter(masks)
        rows = [next(it) if p > 0.0 else torch.ones(B, device=device) for p in drop_probs]
        assert next(it, None) is None
        return torch.stack(rows).float()
    return source


CASES = [
    ("vit_small_patch16_224", dict(num_classes=1000, global_pool="avg", drop_path_rate=0.1), 8, 1.5e-2),
    ("vit_base_patch16_224", dict(num_classes=1000, global_pool="avg", drop_path_rate=0.1), 8, 1.5e-2),
    ("deit_base_distilled_patch16_224", dict(num_classes=1000, drop_path_rate=0.1), 8, 1.5e-2),
    ("vit_large_patch16_384", dict(num_classes=1000, global_pool="avg", drop_path_rate=0.1), 2, 2e-2),
    ("my_vit_mini", dict(num_classes=1000, global_pool="avg", drop_path_rate=0.1), 8, 1.5e-2),
    ("my_vit_xs", dict(num_classes=1000, global_pool="avg", drop_path_rate=0.1), 8, 1.5e-2),   # head_dim 72
]


@pytest.mark.parametrize("name,kw,B,gtol", CASES, ids=[c[0] for c in CASES])
def test_config_activations_logits_and_all_gradients(cuda_device, name, kw, B, gtol):
    from oracle import vit_oracle as O
    from vision_transformers_torch_xla_b200 import ops
    from vision_transformers_torch_xla_b200.losses import SoftTargetCrossEntropy
    from vision_transformers_torch_xla_b200.models import create_model

    dev = cuda_device
    torch.manual_seed(0)
    ref = O.create_model(name, **kw).to(dev)
    mine = create_model(name, **kw).to(dev)
    mine.load_state_dict(ref.state_dict())
    ref.train()
    mine.train()
    distilled = name.startswith("deit_")
    if distilled:
        ref.set_distilled_training(True)
        mine.set_distilled_training(True)
    img = ref.patch_embed.img_size[0]
    g = torch.Generator().manual_seed(3)
    x = torch.randn(B, 3, img, img, generator=g).to(dev)
    tgt = O.mixup_soft_targets(torch.randint(0, 1000, (B,), generator=g)).to(dev)

    acts = {"ref": [], "mine": []}
    grads = {"ref": [], "mine": []}

    def hook(sa, sg):
        def fn(mod, inp, out):
            sa.append(out.detach())
            out.register_hook(lambda gr: sg.append(gr.detach()))
        return fn

    ref_handles = [blk.register_forward_hook(hook(acts["ref"], grads["ref"])) for blk in ref.blocks]
    for blk in mine.blocks:
        blk.register_forward_hook(hook(acts["mine"], grads["mine"]))

    def loss_of(crit, out):
        return crit(out[0], tgt) + crit(out[1], tgt.flip(0)) if distilled else crit(out, tgt)

    rec = record_masks(ref)
    out_ref = ref(x)
    rec.close()
    for h in ref_handles:
        h.remove()
    loss_ref = loss_of(O.SoftTargetCrossEntropy(), out_ref)
    loss_ref.backward()

    # yardstick: the same oracle under torch.autocast(bfloat16) with the same DropPath masks, forward only
    acts["lib"] = []
    lib_hooks = [blk.register_forward_hook(lambda mod, inp, out: acts["lib"].append(out.detach().float())) for blk in ref.blocks]
    mask_it = iter(rec.masks)
    orig_dp = O.DropPath.forward

    def replay_dp(self, t):
        if self.drop_prob == 0.0 or not self.training:
            return t
        return t * next(mask_it).to(t.dtype).view((t.shape[0],) + (1,) * (t.ndim - 1))

    O.DropPath.forward = replay_dp
    try:
        with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
            ref(x)
    finally:
        O.DropPath.forward = orig_dp
        for h in lib_hooks:
            h.remove()
    assert len(rec.masks) == 2 * (len(ref.blocks) - 1)   # dpr[0] = 0 -> block 0 has no DropPath (vision_transformer.py:581)
    assert any(float(m.min()) == 0.0 for m in rec.masks) or B <= 2, "no sample was dropped: the test would be vacuous"

    ops.mask_source = replay(rec.masks)
    try:
        out = mine(x)
    finally:
        ops.mask_source = None
    loss = loss_of(SoftTargetCrossEntropy(), out)
    loss.backward()

    for j, (a, b) in enumerate(zip(out if distilled else (out,), out_ref if distilled else (out_ref,))):
        report(f"{name} logits[{j}]", a, b)
        # (logits: the head's K = D dot products of the pooled features; with two class tokens' heads summed in the
        # distilled model the worst of 8000 elements reaches 2.9e-2 of the RMS)
        assert elem_err(a, b) < 3e-2 and rms_err(a, b) < 1e-2 and rel_err(a, b) < 4e-2
    assert abs(loss.item() - loss_ref.item()) < 2e-3 * abs(loss_ref.item()), (loss.item(), loss_ref.item())
    for i, (a, b) in enumerate(zip(acts["mine"], acts["ref"])):
        if i in (0, len(acts["ref"]) // 2, len(acts["ref"]) - 1):
            report(f"{name} block {i} output", a, b)
    worst_a = max((elem_err(a, b), rel_err(a, b), rms_err(a, b), i) for i, (a, b) in enumerate(zip(acts["mine"], acts["ref"])))
    print(f"[parity] {name}: worst block activation elem {worst_a[0]:.3e} max/rms {worst_a[1]:.3e} rms {worst_a[2]:.3e} "
          f"(block {worst_a[3]})")
    assert elem_err(acts["mine"][0], acts["ref"][0]) < 2e-2
    worst_ratio = 0.0
    for i, (a, b, lib) in enumerate(zip(acts["mine"], acts["ref"], acts["lib"])):
        e, e_lib = elem_err(a, b), elem_err(lib, b)
        worst_ratio = max(worst_ratio, e / e_lib)
        assert e < 3e-2 and rms_err(a, b) < 7e-3 and rel_err(a, b) < 4e-2, i
        assert e <= 1.25 * e_lib + 2e-3, (i, e, e_lib)
    e_lib_last = elem_err(acts["lib"][-1], acts["ref"][-1])
    print(f"[parity] {name}: torch.autocast(bf16) oracle vs fp32 oracle, last block: elem {e_lib_last:.3e} rms "
          f"{rms_err(acts['lib'][-1], acts['ref'][-1]):.3e}; worst ratio (vitk error / library bf16 error) over the blocks {worst_ratio:.2f}")
    worst_g = max((elem_err(a, b), rel_err(a, b), rms_err(a, b), i) for i, (a, b) in enumerate(zip(grads["mine"], grads["ref"])))
    print(f"[parity] {name}: worst residual-stream gradient elem {worst_g[0]:.3e} max/rms {worst_g[1]:.3e} rms {worst_g[2]:.3e} "
          f"(#{worst_g[3]})")
    for a, b in zip(grads["mine"], grads["ref"]):
        assert rms_err(a, b) < 1e-2 and cos_sim(a, b) > 0.9999
        if not distilled:
            assert elem_err(a, b) < 8e-2 and rel_err(a, b) < 1e-1
    refp = dict(ref.named_parameters())
    worst = dict(rms=(0.0, ""), cos=(1.0, ""), mx=(0.0, ""))
    for n, p in mine.named_parameters():
        assert p.grad is not None, n
        gr = refp[n].grad
        if n.endswith("attn.qkv.bias"):
            # the key-bias third has an exactly-zero true gradient (softmax is shift-invariant): compare q and v parts
            D = p.numel() // 3
            a, b = torch.cat([p.grad[:D], p.grad[2 * D:]]), torch.cat([gr[:D], gr[2 * D:]])
        else:
            a, b = p.grad, gr
        r, c, m = rms_err(a, b), cos_sim(a, b), rel_err(a, b)
        worst["rms"] = max(worst["rms"], (r, n))
        worst["cos"] = min(worst["cos"], (c, n))
        worst["mx"] = max(worst["mx"], (m, n))
        assert r < gtol and c > 0.999, (n, r, c)
    print(f"[parity] {name}: parameter gradients: worst rms {worst['rms'][0]:.3e} ({worst['rms'][1]}), worst cosine "
          f"{worst['cos'][0]:.6f} ({worst['cos'][1]}), worst max/rms {worst['mx'][0]:.3e} ({worst['mx'][1]})")


def test_droppath_kernel_statistics_and_reproducibility(cuda_device):
    """vitk_droppath_masks: values are 0 or 1/keep, the keep rate matches, rows with p = 0 are all ones, the same
    (seed, offset) gives the same masks and the next call different ones."""
    from vision_transformers_torch_xla_b200 import _lib as L

    probs = [0.0, 0.1, 0.25, 0.5, 1.0]
    B = 20000
    a = torch.empty(len(probs), B, device=cuda_device)
    b = torch.empty_like(a)
    c = torch.empty_like(a)
    L.droppath_masks(a, probs, 1234, 7)
    L.droppath_masks(b, probs, 1234, 7)
    L.droppath_masks(c, probs, 1234, 8)
    assert torch.equal(a, b) and not torch.equal(a[1:4], c[1:4])
    assert torch.equal(a[0], torch.ones(B, device=cuda_device)) and float(a[4].abs().max()) == 0.0
    for i, p in enumerate(probs[1:4], start=1):
        keep = 1.0 - p
        vals = torch.unique(a[i])
        assert vals.numel() == 2 and float(vals[0]) == 0.0 and abs(float(vals[1]) - 1.0 / keep) < 1e-6
        rate = float((a[i] > 0).float().mean())
        assert abs(rate - keep) < 4 * (keep * p / B) ** 0.5 + 1e-3, (p, rate)
        assert abs(float(a[i].mean()) - 1.0) < 0.03   # unbiased: E[mask / keep] = 1
    # rows are independent draws
    assert abs(float(((a[2] > 0) & (a[3] > 0)).float().mean()) - 0.75 * 0.5) < 0.02


def test_model_draws_fresh_masks_each_forward(cuda_device):
    from vision_transformers_torch_xla_b200 import ops
    from vision_transformers_torch_xla_b200.models import create_model

    torch.manual_seed(0)
    m = create_model("vit_tiny_patch16_224", num_classes=10, global_pool="avg", drop_path_rate=0.5).to(cuda_device)
    m.train()
    x = torch.randn(16, 3, 224, 224, device=cuda_device)
    with torch.no_grad():
        a, b = m(x), m(x)
        assert not torch.equal(a, b), "two training forwards used the same DropPath masks"
        m.eval()
        c, d = m(x), m(x)
        assert torch.equal(c, d)
    n0 = ops._mask_calls
    m.train()
    with torch.no_grad():
        m(x)
    assert ops._mask_calls == n0 + 1, "all 24 masks of a forward pass come from ONE launch"


def test_clip_grad_norm_matches_torch(cuda_device):
    """engine.clip_grad_norm_ + FusedAdamW.step against torch.nn.utils.clip_grad_norm_ + torch.optim.AdamW fed the
    SAME gradients: the returned norm, and the weights after the step (the clip coefficient never leaves the device:
    the AdamW kernel multiplies it in)."""
    from oracle import vit_oracle as O
    from vision_transformers_torch_xla_b200 import engine, optim_factory
    from vision_transformers_torch_xla_b200.losses import SoftTargetCrossEntropy
    from vision_transformers_torch_xla_b200.models import create_model

    dev = cuda_device
    torch.manual_seed(0)
    ref = O.create_model("vit_tiny_patch16_224", num_classes=1000, global_pool="avg").to(dev)
    mine = create_model("vit_tiny_patch16_224", num_classes=1000, global_pool="avg").to(dev)
    mine.load_state_dict(ref.state_dict())
    mine.train()

    class Args:
        opt, lr, weight_decay, opt_eps, opt_betas = "adamw", 1e-3, 0.05, 1e-8, None

    opt = optim_factory.create_optimizer(Args, mine)
    opt_ref = O.create_optimizer(ref, lr=1e-3, weight_decay=0.05)
    x = torch.randn(4, 3, 224, 224, device=dev)
    y = O.mixup_soft_targets(torch.randint(0, 1000, (4,))).to(dev)
    for max_norm in (0.05, 1e4):   # clipping active / inactive
        SoftTargetCrossEntropy()(mine(x), y).backward()
        refp = dict(ref.named_parameters())
        for n, p in mine.named_parameters():
            refp[n].grad = p.grad.detach().clone()
        want_norm = torch.nn.utils.clip_grad_norm_(ref.parameters(), max_norm)
        got_norm = engine.clip_grad_norm_(opt, max_norm)
        assert abs(float(got_norm) - float(want_norm)) < 1e-4 * float(want_norm), (float(got_norm), float(want_norm))
        opt.step()
        opt.zero_grad()
        opt_ref.step()
        opt_ref.zero_grad()
        for n, p in mine.named_parameters():
            assert rms_err(p.data, refp[n].data) < 1e-5, (max_norm, n, rms_err(p.data, refp[n].data))
        assert opt.grad_scale_dev is None, "the clip coefficient applies to one step only"


def test_optimizer_state_loads_before_first_forward_and_survives_a_plan_rebuild(cuda_device, tmp_path):
    """The reference's resume order — create model, create optimizer, auto_load_model, THEN the first forward
    (/root/reference/main.py:979) — and a plan rebuild (reset_classifier) keeping the Adam moments by parameter."""
    import argparse

    from oracle import vit_oracle as O
    from vision_transformers_torch_xla_b200 import optim_factory, utils
    from vision_transformers_torch_xla_b200.losses import SoftTargetCrossEntropy
    from vision_transformers_torch_xla_b200.models import create_model

    class OptArgs:
        opt, lr, weight_decay, opt_eps, opt_betas = "adamw", 1e-3, 0.05, 1e-8, None

    dev = cuda_device
    x = torch.randn(4, 3, 224, 224, device=dev)
    y = O.mixup_soft_targets(torch.randint(0, 1000, (4,))).to(dev)
    crit = SoftTargetCrossEntropy()

    def make():
        torch.manual_seed(0)
        m = create_model("vit_tiny_patch16_224", num_classes=1000, global_pool="avg").to(dev)
        return m, optim_factory.create_optimizer(OptArgs, m)

    a, oa = make()
    for _ in range(2):
        crit(a(x), y).backward()
        oa.step()
        oa.zero_grad()
    args = argparse.Namespace(output_dir=str(tmp_path), save_ckpt_num=3, save_ckpt_freq=1, auto_resume=True, resume="",
                              start_epoch=0)
    utils.save_model(args, 0, a, a, oa, None)
    b, ob = make()
    utils.auto_load_model(args, b, b, ob, None)           # no forward has run on b yet
    assert ob._step == 2
    for (n, pa), (_, pb) in zip(a.named_parameters(), b.named_parameters()):
        assert torch.equal(pa.data, pb.data), n
        assert torch.equal(oa.state[pa]["exp_avg"], ob.state[pb]["exp_avg"]), n
        assert torch.equal(oa.state[pa]["exp_avg_sq"], ob.state[pb]["exp_avg_sq"]), n
    # a plan rebuild (new head -> new flat store) carries the moments of the surviving parameters over
    m_before = ob.state[b.blocks[3].mlp.fc1.weight]["exp_avg"].clone()
    old_head = {id(b.head.weight), id(b.head.bias)}
    b.reset_classifier(10)
    for gp in ob.param_groups:   # swap the old head's parameters for the new ones (weight -> decay group, bias -> no_decay)
        had_w = any(id(p) in old_head and p.ndim == 2 for p in gp["params"])
        had_b = any(id(p) in old_head and p.ndim == 1 for p in gp["params"])
        gp["params"] = [p for p in gp["params"] if id(p) not in old_head] + ([b.head.weight] if had_w else []) + \
            ([b.head.bias] if had_b else [])
    y10 = torch.softmax(torch.randn(4, 10, device=dev), -1)
    crit(b(x), y10).backward()
    ob.step()
    assert ob._step == 3
    m_after = ob.state[b.blocks[3].mlp.fc1.weight]["exp_avg"]
    assert float(m_before.abs().max()) > 0
    # m_after = 0.9 * m_before + 0.1 * g: had the moments been reset it would be 0.1 * g, uncorrelated with m_before
    assert cos_sim(m_after, m_before) > 0.5


def test_grad_checkpointing_gives_the_same_step_with_less_memory(cuda_device):
    """set_grad_checkpointing(True) (/root/reference/models/vision_transformer.py:686-694, 945-946): every block keeps only
    its input and re-runs its forward kernels in the backward, with the DropPath / dropout masks of the forward pass.  Same
    logits bit for bit, the same gradients up to the summation order of the split-K weight gradients, a smaller peak."""
    from vision_transformers_torch_xla_b200 import ops
    from vision_transformers_torch_xla_b200.losses import SoftTargetCrossEntropy
    from vision_transformers_torch_xla_b200.models import create_model

    dev, B = cuda_device, 16
    torch.manual_seed(0)
    model = create_model("vit_small_patch16_224", num_classes=1000, global_pool="avg", drop_path_rate=0.1, proj_drop_rate=0.1).to(dev)
    model.train()
    x = torch.randn(B, 3, 224, 224, device=dev)
    tgt = torch.softmax(torch.randn(B, 1000, device=dev), -1)
    crit = SoftTargetCrossEntropy()
    masks, drops = {}, {}

    def dp_source(drop_probs, nb, device):      # the same DropPath factors and dropout masks in both passes
        key = tuple(drop_probs)
        if key not in masks:
            g = torch.Generator().manual_seed(5)
            masks[key] = torch.stack([(torch.rand(nb, generator=g) >= p).float() / (1 - p) for p in drop_probs]).to(device)
        return masks[key]

    def do_source(site, rows, cols, p, device):
        if site not in drops:
            g = torch.Generator().manual_seed(hash(site) % 2 ** 31)
            drops[site] = (torch.rand(rows, cols, generator=g) >= p).to(torch.uint8).to(device)
        return drops[site]

    def run(ckpt):
        model.set_grad_checkpointing(ckpt)
        model.zero_grad(set_to_none=True)
        torch.cuda.synchronize()
        torch.cuda.reset_peak_memory_stats()
        base = torch.cuda.memory_allocated()
        ops.mask_source, ops.dropout_source = dp_source, do_source
        try:
            out = model(x)
            crit(out, tgt).backward()
        finally:
            ops.mask_source, ops.dropout_source = None, None
        torch.cuda.synchronize()
        peak = torch.cuda.max_memory_allocated() - base
        return out.detach().clone(), {n: p.grad.detach().clone() for n, p in model.named_parameters()}, peak

    out_a, grads_a, peak_a = run(False)
    out_b, grads_b, peak_b = run(True)
    model.set_grad_checkpointing(False)
    assert torch.equal(out_a, out_b)
    for n in grads_a:
        assert rms_err(grads_b[n], grads_a[n]) < 1e-5, (n, rms_err(grads_b[n], grads_a[n]))
    print(f"[ckpt] peak activation memory {peak_a / 2 ** 20:.0f} MiB -> {peak_b / 2 ** 20:.0f} MiB")
    assert peak_b < 0.5 * peak_a
