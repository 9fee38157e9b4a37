"""CPU tests of the oracle (oracle/vit_oracle.py).

The reference pins no numerical results ("parity unpinned"), so the oracle is pinned by the
structural known-answers the survey lists (parameter counts, state_dict layout, shapes, init loss, closed-form
schedules, the reference quirks), by an INDEPENDENT implementation of the same architecture
(torchvision.models.VisionTransformer) on shared weights, and by committed golden vectors (regression pin).
"""
import math
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import vit_oracle as O

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "vit_micro.pt")


@pytest.mark.parametrize("name,kw,count", [
    ("vit_tiny_patch16_224", dict(global_pool="token"), 5_717_416),
    ("vit_tiny_patch16_224", dict(global_pool="avg"), 5_717_416),
    ("vit_small_patch16_224", dict(global_pool="avg"), 22_050_664),
    ("vit_base_patch16_224", dict(global_pool="avg"), 86_567_656),
    ("deit_base_distilled_patch16_224", dict(), 87_338_192),
    ("vit_large_patch16_384", dict(global_pool="avg"), 304_715_752),
])
def test_param_counts(name, kw, count):
    with torch.device("meta"):
        m = O.create_model(name, num_classes=1000, **kw)
    assert sum(p.numel() for p in m.parameters()) == count


def test_state_dict_layout_appendix_b():
    with torch.device("meta"):
        m = O.create_model("vit_base_patch16_224", num_classes=1000, global_pool="avg")
        t = O.create_model("vit_base_patch16_224", num_classes=1000, global_pool="token")
        d = O.create_model("deit_base_distilled_patch16_224", num_classes=1000)
        l384 = O.create_model("vit_large_patch16_384", num_classes=1000, global_pool="avg")
    sd = {k: tuple(v.shape) for k, v in m.state_dict().items()}
    assert len(list(m.parameters())) == 152  # 4 + 12*12 + 2 + 2
    assert sd["cls_token"] == (1, 1, 768) and sd["pos_embed"] == (1, 197, 768)
    assert sd["patch_embed.proj.weight"] == (768, 3, 16, 16) and sd["patch_embed.proj.bias"] == (768,)
    assert sd["blocks.0.attn.qkv.weight"] == (2304, 768) and sd["blocks.0.attn.qkv.bias"] == (2304,)
    assert sd["blocks.11.mlp.fc1.weight"] == (3072, 768) and sd["blocks.11.mlp.fc2.weight"] == (768, 3072)
    assert "fc_norm.weight" in sd and "norm.weight" not in sd
    assert "norm.weight" in t.state_dict() and "fc_norm.weight" not in t.state_dict()
    dsd = d.state_dict()
    assert tuple(dsd["dist_token"].shape) == (1, 1, 768) and tuple(dsd["pos_embed"].shape) == (1, 198, 768)
    assert "head_dist.weight" in dsd
    assert tuple(l384.pos_embed.shape) == (1, 577, 1024)


def test_init_statistics_and_loss_at_init():
    torch.manual_seed(0)
    m = O.create_model("vit_tiny_patch16_224", num_classes=1000, global_pool="avg")
    assert float(m.cls_token.abs().max()) < 1e-4  # normal_(std=1e-6)
    assert abs(float(m.pos_embed.std()) - 0.02) < 2e-3
    assert abs(float(m.blocks[3].mlp.fc1.weight.std()) - 0.02) < 1e-3
    assert float(m.blocks[3].mlp.fc1.bias.abs().max()) == 0.0
    assert float(m.fc_norm.weight.min()) == 1.0
    m.eval()
    x = torch.randn(4, 3, 224, 224)
    with torch.no_grad():
        logits = m(x)
    assert logits.shape == (4, 1000)
    loss = F.cross_entropy(logits, torch.randint(0, 1000, (4,)))
    assert abs(loss.item() - math.log(1000)) < 0.15  # ~ ln(1000) = 6.9078 at init


def test_fused_and_unfused_attention_agree():
    torch.manual_seed(0)
    a = O.Attention(192, num_heads=3, qkv_bias=True, fused=True)
    b = O.Attention(192, num_heads=3, qkv_bias=True, fused=False)
    b.load_state_dict(a.state_dict())
    x = torch.randn(2, 197, 192)
    assert torch.allclose(a(x), b(x), atol=1e-5, rtol=1e-4)


def test_against_torchvision_vit_independent_implementation():
    """Same architecture, written independently (torchvision): shared weights must give the same logits."""
    tv = pytest.importorskip("torchvision.models.vision_transformer")
    torch.manual_seed(0)
    D, depth, heads = 128, 2, 2
    ref = tv.VisionTransformer(image_size=64, patch_size=16, num_layers=depth, num_heads=heads, hidden_dim=D,
                               mlp_dim=4 * D, num_classes=10)
    m = O.VisionTransformer(img_size=64, patch_size=16, embed_dim=D, depth=depth, num_heads=heads, num_classes=10,
                            global_pool="token")
    sd = {}
    sd["conv_proj.weight"], sd["conv_proj.bias"] = m.patch_embed.proj.weight, m.patch_embed.proj.bias
    sd["class_token"], sd["encoder.pos_embedding"] = m.cls_token, m.pos_embed
    for i, blk in enumerate(m.blocks):
        p = f"encoder.layers.encoder_layer_{i}."
        sd[p + "ln_1.weight"], sd[p + "ln_1.bias"] = blk.norm1.weight, blk.norm1.bias
        sd[p + "self_attention.in_proj_weight"], sd[p + "self_attention.in_proj_bias"] = blk.attn.qkv.weight, blk.attn.qkv.bias
        sd[p + "self_attention.out_proj.weight"], sd[p + "self_attention.out_proj.bias"] = blk.attn.proj.weight, blk.attn.proj.bias
        sd[p + "ln_2.weight"], sd[p + "ln_2.bias"] = blk.norm2.weight, blk.norm2.bias
        sd[p + "mlp.0.weight"], sd[p + "mlp.0.bias"] = blk.mlp.fc1.weight, blk.mlp.fc1.bias
        sd[p + "mlp.3.weight"], sd[p + "mlp.3.bias"] = blk.mlp.fc2.weight, blk.mlp.fc2.bias
    sd["encoder.ln.weight"], sd["encoder.ln.bias"] = m.norm.weight, m.norm.bias
    sd["heads.head.weight"], sd["heads.head.bias"] = m.head.weight, m.head.bias
    with torch.no_grad():
        for k in sd:
            if k.endswith("bias") or ".ln" in k:
                sd[k].add_(0.1 * torch.randn_like(sd[k]))
    ref.load_state_dict({k: v.detach().clone() for k, v in sd.items()}, strict=True)
    ref.eval()
    m.eval()
    x = torch.randn(3, 3, 64, 64)
    with torch.no_grad():
        assert torch.allclose(m(x), ref(x), atol=2e-5, rtol=1e-4)


def test_against_huggingface_vit_second_independent_implementation():
    """A second independently written ViT (transformers.ViTForImageClassification, the architecture timm's checkpoints
    were ported to): shared weights must give the same logits AND the same parameter gradients."""
    tr = pytest.importorskip("transformers")
    torch.manual_seed(0)
    D, depth, heads = 128, 2, 2
    cfg = tr.ViTConfig(hidden_size=D, num_hidden_layers=depth, num_attention_heads=heads, intermediate_size=4 * D,
                       hidden_act="gelu", layer_norm_eps=1e-6, image_size=64, patch_size=16, num_channels=3, qkv_bias=True,
                       num_labels=10, hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0)
    ref = tr.ViTForImageClassification(cfg)
    m = O.VisionTransformer(img_size=64, patch_size=16, embed_dim=D, depth=depth, num_heads=heads, num_classes=10,
                            global_pool="token")
    with torch.no_grad():
        for prm in m.parameters():
            if prm.ndim == 1:
                prm.add_(0.1 * torch.randn_like(prm))
    sd = {"vit.embeddings.cls_token": m.cls_token, "vit.embeddings.position_embeddings": m.pos_embed,
          "vit.embeddings.patch_embeddings.projection.weight": m.patch_embed.proj.weight,
          "vit.embeddings.patch_embeddings.projection.bias": m.patch_embed.proj.bias,
          "vit.layernorm.weight": m.norm.weight, "vit.layernorm.bias": m.norm.bias,
          "classifier.weight": m.head.weight, "classifier.bias": m.head.bias}
    for i, blk in enumerate(m.blocks):
        p = f"vit.encoder.layer.{i}."
        wq, wk, wv = blk.attn.qkv.weight.chunk(3, 0)
        bq, bk, bv = blk.attn.qkv.bias.chunk(3, 0)
        for nm, w, b in (("query", wq, bq), ("key", wk, bk), ("value", wv, bv)):
            sd[p + f"attention.attention.{nm}.weight"], sd[p + f"attention.attention.{nm}.bias"] = w, b
        sd[p + "attention.output.dense.weight"], sd[p + "attention.output.dense.bias"] = blk.attn.proj.weight, blk.attn.proj.bias
        sd[p + "layernorm_before.weight"], sd[p + "layernorm_before.bias"] = blk.norm1.weight, blk.norm1.bias
        sd[p + "layernorm_after.weight"], sd[p + "layernorm_after.bias"] = blk.norm2.weight, blk.norm2.bias
        sd[p + "intermediate.dense.weight"], sd[p + "intermediate.dense.bias"] = blk.mlp.fc1.weight, blk.mlp.fc1.bias
        sd[p + "output.dense.weight"], sd[p + "output.dense.bias"] = blk.mlp.fc2.weight, blk.mlp.fc2.bias
    ref.load_state_dict({k: v.detach().clone() for k, v in sd.items()}, strict=True)
    ref.train()
    m.train()
    x = torch.randn(3, 3, 64, 64)
    t = torch.softmax(torch.randn(3, 10), -1)
    out_ref = ref(pixel_values=x).logits
    out = m(x)
    assert torch.allclose(out, out_ref, atol=2e-5, rtol=1e-4)
    O.SoftTargetCrossEntropy()(out, t).backward()
    O.SoftTargetCrossEntropy()(out_ref, t).backward()
    g_ref = dict(ref.named_parameters())
    assert torch.allclose(m.blocks[0].mlp.fc1.weight.grad, g_ref["vit.encoder.layer.0.intermediate.dense.weight"].grad, atol=1e-6, rtol=1e-3)
    assert torch.allclose(m.blocks[1].attn.qkv.weight.grad[:D], g_ref["vit.encoder.layer.1.attention.attention.query.weight"].grad,
                          atol=1e-6, rtol=1e-3)
    assert torch.allclose(m.patch_embed.proj.weight.grad, g_ref["vit.embeddings.patch_embeddings.projection.weight"].grad,
                          atol=1e-6, rtol=1e-3)


def test_losses_closed_forms():
    torch.manual_seed(0)
    x = torch.randn(5, 11)
    y = torch.randint(0, 11, (5,))
    assert torch.allclose(O.LabelSmoothingCrossEntropy(0.1)(x, y), F.cross_entropy(x, y, label_smoothing=0.1), atol=1e-6)
    t = torch.softmax(torch.randn(5, 11), -1)
    assert torch.allclose(O.SoftTargetCrossEntropy()(x, t), F.cross_entropy(x, t), atol=1e-6)
    z = torch.randn(5, 11)
    kd = O.DistillationLoss(O.LabelSmoothingCrossEntropy(0.1), alpha=0.7, temperature=4.0)
    pt = torch.softmax(z / 4, 1)
    manual = 0.3 * F.cross_entropy(x, y, label_smoothing=0.1) + 0.7 * 16.0 * (
        pt * (pt.log() - torch.log_softmax(x / 4, 1))).sum() / 5
    assert torch.allclose(kd((x, z), y), manual, atol=1e-5)
    assert torch.allclose(kd(x, y), F.cross_entropy(x, y, label_smoothing=0.1), atol=1e-6)  # tensor -> base only


def test_student_wrapper_contract_like_test_kd():
    """/root/reference/test_kd.py:106-124: train -> (student, teacher) tuple, eval -> tensor."""
    s = O.VisionTransformer(img_size=32, embed_dim=64, depth=1, num_heads=1, num_classes=7, global_pool="avg")
    t = O.VisionTransformer(img_size=32, embed_dim=64, depth=1, num_heads=1, num_classes=7, global_pool="avg")
    w = O.StudentWithDistillation(s, t)
    x = torch.randn(4, 3, 32, 32)
    w.train()
    out = w(x)
    assert isinstance(out, tuple) and out[0].shape == (4, 7) and not out[1].requires_grad
    w.eval()
    assert isinstance(w(x), torch.Tensor)


def test_cosine_scheduler_closed_form():
    s = O.cosine_scheduler(4e-3, 1e-6, epochs=3, niter_per_ep=10, warmup_epochs=1)
    assert len(s) == 30 and s[0] == 0.0 and abs(s[9] - 4e-3) < 1e-12
    i = 7
    want = 1e-6 + 0.5 * (4e-3 - 1e-6) * (1 + math.cos(math.pi * i / 20))
    assert abs(s[10 + i] - want) < 1e-12
    # reference quirk (utils/__init__.py:671-676): warmup_steps shortens the cosine part but the warm-up ramp is
    # only generated when warmup_epochs > 0, so warmup_steps alone trips the function's own length assert
    with pytest.raises(AssertionError):
        O.cosine_scheduler(1.0, 0.0, 2, 5, warmup_epochs=0, warmup_steps=3)
    s2 = O.cosine_scheduler(1.0, 0.0, 2, 5, warmup_epochs=1, warmup_steps=3)
    assert len(s2) == 10 and s2[2] == 1.0


def test_param_groups_and_wd_quirk():
    with torch.device("meta"):
        m = O.create_model("vit_tiny_patch16_224", num_classes=1000, global_pool="avg")
    groups = O.get_parameter_groups(m, 0.05, m.no_weight_decay())
    by_wd = {g["weight_decay"]: g for g in groups}
    assert len(groups) == 2 and set(by_wd) == {0.0, 0.05}
    # decay: patch proj + 4 matrices per block + head = 1 + 48 + 1
    assert len(by_wd[0.05]["params"]) == 50 and len(by_wd[0.0]["params"]) == 102
    assert all(g["lr_scale"] == 1.0 for g in groups)

    class FakeOpt:
        param_groups = [dict(weight_decay=0.05, lr_scale=1.0), dict(weight_decay=0.0, lr_scale=1.0)]

    O.apply_schedules(FakeOpt, 0, np.array([1e-3]), np.array([0.04]), wd_quirk=True)
    assert [g["weight_decay"] for g in FakeOpt.param_groups] == [0.04, 0.04]  # engine.py:102 as written
    FakeOpt.param_groups[1]["weight_decay"] = 0.0
    O.apply_schedules(FakeOpt, 0, np.array([1e-3]), np.array([0.03]), wd_quirk=False)
    assert [g["weight_decay"] for g in FakeOpt.param_groups] == [0.03, 0.0]  # the commented-out original (> 0)


def test_train_step_learns_and_grad_accumulates():
    torch.manual_seed(0)
    m = O.VisionTransformer(img_size=32, embed_dim=64, depth=2, num_heads=1, num_classes=10, global_pool="avg")
    opt = O.create_optimizer(m, lr=1e-3, weight_decay=0.05)
    x = torch.randn(8, 3, 32, 32)
    y = O.mixup_soft_targets(torch.randint(0, 10, (8,)), 10)
    assert torch.allclose(y.sum(1), torch.ones(8), atol=1e-6)
    first = None
    for _ in range(30):
        loss, _ = O.train_step(m, O.SoftTargetCrossEntropy(), opt, x, y)
        first = first if first is not None else float(loss)
    assert float(loss) < 0.7 * first
    stats = O.train_one_epoch(m, O.SoftTargetCrossEntropy(), [(x, y)] * 4, opt, update_freq=2,
                              lr_schedule_values=O.cosine_scheduler(1e-3, 0, 1, 2), wd_schedule_values=None,
                              num_training_steps_per_epoch=2)
    assert math.isfinite(stats["loss"])


@pytest.mark.parametrize("pool", ["avg", "token"])
def test_golden_vectors(pool):
    """Committed fixture (tests/golden/make_golden.py): the oracle must keep reproducing it bit-for-bit-ish."""
    g = torch.load(GOLDEN)
    case = g[pool]
    torch.set_num_threads(1)
    m = O.VisionTransformer(**case["cfg"])
    sd = dict(g["state_dict_avg"])
    if pool == "token":
        sd = {(k.replace("fc_norm.", "norm.")): v for k, v in sd.items()}
    m.load_state_dict(sd)
    m.train()
    logits = m(case["x"])
    loss = O.SoftTargetCrossEntropy()(logits, case["target"])
    loss.backward()
    assert torch.allclose(logits, case["logits"], atol=1e-5, rtol=1e-5)
    assert abs(float(loss) - float(case["loss"])) < 1e-6
    named = dict(m.named_parameters())
    for n, gr in case["grads"].items():
        assert torch.allclose(named[n].grad, gr, atol=1e-6, rtol=1e-4), n
