"""Host-side logic on CPU: registry / factory, module surface and state_dict layout against the oracle,
parameter-group rule, schedules, meters — and that the product refuses to compute without its CUDA kernels."""

import numpy as np
import pytest
import torch

from oracle import vit_oracle as O
from vision_transformers_torch_xla_b200 import losses, optim_factory, utils
from vision_transformers_torch_xla_b200._lib import VitkError
from vision_transformers_torch_xla_b200.models import (Attention, Block, Mlp, PatchEmbed, VisionTransformer,
                                                     VisionTransformerDistilled, create_model, is_model, list_models)


def test_registry_lists_the_reference_entrypoints():
    names = list_models()
    for n in ("vit_tiny_patch16_224", "vit_small_patch16_224", "vit_base_patch16_224", "vit_large_patch16_384",
              "deit_base_distilled_patch16_224", "my_vit_ti", "my_vit_s", "my_vit_b", "my_vit_l"):
        assert n in names and is_model(n)
    assert list_models("my_vit_*") == ["my_vit_b", "my_vit_l", "my_vit_mini", "my_vit_s", "my_vit_ti", "my_vit_xs"]


@pytest.mark.parametrize("name,kw", [
    ("vit_tiny_patch16_224", dict(num_classes=1000, global_pool="avg", drop_path_rate=0.1)),
    ("vit_tiny_patch16_224", dict(num_classes=1000, global_pool="token")),
    ("vit_base_patch16_224", dict(num_classes=1000, global_pool="avg")),
    ("deit_base_distilled_patch16_224", dict(num_classes=1000)),
    ("vit_large_patch16_384", dict(num_classes=1000, global_pool="avg")),
    ("my_vit_b", dict(num_classes=1000, global_pool="avg")),
])
def test_state_dict_layout_matches_oracle(name, kw):
    with torch.device("meta"):
        mine = create_model(name, pretrained=False, **kw)
        oname = {"my_vit_b": "vit_base_patch16_224"}.get(name, name)
        ref = O.create_model(oname, **kw)
    a = {k: tuple(v.shape) for k, v in mine.state_dict().items()}
    b = {k: tuple(v.shape) for k, v in ref.state_dict().items()}
    assert a == b
    assert [n for n, _ in mine.named_parameters()] == [n for n, _ in ref.named_parameters()]


def test_create_model_kwargs_contract():
    m = create_model("vit_tiny_patch16_224", pretrained=False, num_classes=10, global_pool="avg", drop_path_rate=None,
                     img_size=None)  # None kwargs are pruned (reference _factory.py:108)
    assert m.num_classes == 10 and m.global_pool == "avg" and m.patch_embed.img_size == (224, 224)
    assert isinstance(m.fc_norm, torch.nn.LayerNorm) and isinstance(m.norm, torch.nn.Identity)
    assert m.fc_norm.eps == 1e-6 and m.blocks[0].norm1.eps == 1e-6
    assert m.no_weight_decay() == {"pos_embed", "cls_token", "dist_token"}
    assert m.get_classifier() is m.head
    m2 = create_model("vit_tiny_patch16_224", drop_path_rate=0.2)
    dps = [b.drop_path1.drop_prob if hasattr(b.drop_path1, "drop_prob") else 0.0 for b in m2.blocks]
    assert dps[0] == 0.0 and abs(dps[-1] - 0.2) < 1e-6 and dps == sorted(dps)  # linspace rule (:581)
    d = create_model("deit_tiny_distilled_patch16_224")
    assert isinstance(d, VisionTransformerDistilled) and d.num_prefix_tokens == 2 and d.distilled_training is False
    assert create_model("vit_large_patch16_384").pos_embed.shape == (1, 577, 1024)
    with pytest.raises(RuntimeError, match="Unknown model"):
        create_model("convnext_tiny")
    with pytest.raises(NotImplementedError):
        create_model("vit_tiny_patch16_224", pretrained=True)


def test_options_outside_the_fast_path_raise_instead_of_falling_back():
    for kw in (dict(qk_norm=True), dict(reg_tokens=4), dict(pre_norm=True), dict(global_pool="map"),
               dict(dynamic_img_size=True), dict(patch_drop_rate=0.1), dict(no_embed_class=True)):
        with pytest.raises(NotImplementedError):
            VisionTransformer(embed_dim=64, depth=1, num_heads=1, **kw)
    with pytest.raises(NotImplementedError):
        Attention(384, num_heads=4)  # head_dim 96 is wider than a 64-wide head tile plus its 16-column tail
    assert Attention(144, num_heads=3).head_dim == 48   # narrower heads run zero-padded (my_vit_mini)
    assert Attention(288, num_heads=4).head_dim == 72   # my_vit_xs: a second, zero-padded tile per operand
    with pytest.raises(NotImplementedError):
        Mlp(64, act_layer=torch.nn.ReLU)
    with pytest.raises(NotImplementedError):
        Block(64, 1)(torch.zeros(1, 5, 64), attn_mask=torch.zeros(5, 5))


def test_product_refuses_to_run_on_cpu():
    """No CPU fallback: a CPU model / CPU tensors raise VitkError instead of silently using PyTorch kernels."""
    m = VisionTransformer(img_size=32, embed_dim=64, depth=1, num_heads=1, num_classes=8, global_pool="avg")
    with pytest.raises(VitkError):
        m(torch.zeros(1, 3, 32, 32))
    with pytest.raises(VitkError):
        PatchEmbed(32, 16, 3, 64)(torch.zeros(1, 3, 32, 32))
    with pytest.raises(VitkError):
        losses.SoftTargetCrossEntropy()(torch.zeros(2, 8), torch.full((2, 8), 0.125))
    with pytest.raises(AssertionError):
        m.patch_embed._check(torch.zeros(1, 3, 48, 48))  # strict input size, as timm's PatchEmbed


def test_parameter_groups_match_oracle_rule():
    with torch.device("meta"):
        m = create_model("vit_tiny_patch16_224", num_classes=1000, global_pool="avg")
        ref = O.create_model("vit_tiny_patch16_224", num_classes=1000, global_pool="avg")
    mine = optim_factory.get_parameter_groups(m, 0.05, m.no_weight_decay())
    want = O.get_parameter_groups(ref, 0.05, ref.no_weight_decay())
    assert [(g["weight_decay"], g["lr_scale"], len(g["params"])) for g in mine] == \
           [(g["weight_decay"], g["lr_scale"], len(g["params"])) for g in want]
    # layer-decay hooks keep the reference's group naming contract
    groups = optim_factory.get_parameter_groups(m, 0.05, (), get_num_layer=lambda n: 0 if "blocks" not in n else 1,
                                                get_layer_scale=lambda i: 0.5 ** i)
    assert sorted({g["lr_scale"] for g in groups}) == [0.5, 1.0]


def test_create_optimizer_contract():
    m = VisionTransformer(img_size=32, embed_dim=64, depth=1, num_heads=1, num_classes=8, global_pool="avg")

    class Args:
        opt, lr, weight_decay, opt_eps, opt_betas = "adamw", 4e-3, 0.05, 1e-8, (0.9, 0.95)

    opt = optim_factory.create_optimizer(Args, m)
    assert isinstance(opt, torch.optim.Optimizer) and len(opt.param_groups) == 2
    assert {g["weight_decay"] for g in opt.param_groups} == {0.0, 0.05}
    assert all(g["lr"] == 4e-3 and g["lr_scale"] == 1.0 and g["betas"] == (0.9, 0.95) for g in opt.param_groups)
    with pytest.raises(VitkError):
        opt.step()  # parameters are not in a CUDA ParamStore: loud failure, no eager AdamW
    Args.opt = "sgd"
    with pytest.raises(NotImplementedError):
        optim_factory.create_optimizer(Args, m)


def test_cosine_scheduler_equals_oracle():
    a = utils.cosine_scheduler(4e-3, 1e-6, 5, 7, warmup_epochs=2, start_warmup_value=1e-5)
    b = O.cosine_scheduler(4e-3, 1e-6, 5, 7, warmup_epochs=2, start_warmup_value=1e-5)
    assert np.array_equal(a, b) and len(a) == 35


def test_meters():
    ml = utils.MetricLogger()
    for v in (1.0, 2.0, 6.0):
        ml.update(loss=v, lr=None)
    assert ml.meters["loss"].global_avg == 3.0 and ml.meters["loss"].median == 2.0 and "lr" not in ml.meters
    ml.synchronize_between_processes()  # no process group: no-op
    out = torch.tensor([[0.1, 0.9, 0.0], [0.8, 0.1, 0.1]])
    acc1, acc2 = utils.accuracy(out, torch.tensor([1, 2]), topk=(1, 2))
    assert acc1.item() == 50.0 and acc2.item() == 50.0


def test_distillation_wrapper_contract_on_cpu_models():
    class Tiny(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.l = torch.nn.Linear(4, 3)

        def forward(self, x):
            return self.l(x)

    w = losses.StudentWithDistillation(Tiny(), Tiny())
    x = torch.randn(2, 4)
    w.train()
    s, t = w(x)
    assert s.requires_grad and not t.requires_grad
    w.eval()
    assert isinstance(w(x), torch.Tensor)
    crit = losses.DistillationLoss(torch.nn.MSELoss(), 0.7, 4.0)
    with pytest.raises(NotImplementedError):
        crit((s, t), torch.zeros(2, 3))  # base criterion the fused kernel cannot absorb


def test_dropout_and_checkpointing_options_are_accepted_on_the_host_side():
    """The nn.Dropout sites of the reference (pos_drop, attn_drop, proj_drop, Mlp.drop1 / drop2, head_drop) and gradient
    checkpointing are built (DESIGN.md sections 4.7 / 4.8)."""
    from vision_transformers_torch_xla_b200 import ops

    m = VisionTransformer(embed_dim=64, depth=2, num_heads=1, drop_rate=0.1, pos_drop_rate=0.2, proj_drop_rate=0.3,
                          attn_drop_rate=0.4)
    assert m.head_drop.p == 0.1 and m.pos_drop.p == 0.2 and all(b.attn.attn_drop.p == 0.4 for b in m.blocks)
    assert all(b.attn.proj_drop.p == 0.3 and b.mlp.drop1.p == 0.3 and b.mlp.drop2.p == 0.3 for b in m.blocks)
    mlp = Mlp(64, 128, drop=(0.1, 0.2))
    assert (mlp.drop1.p, mlp.drop2.p) == (0.1, 0.2)
    assert Attention(64, num_heads=1, attn_drop=0.1).attn_drop.p == 0.1
    assert ops.dropout_keep_mask("site", 4, 8, 0.0, "cpu") is None          # p = 0: no mask, no kernel
    with pytest.raises(ValueError):
        ops.dropout_keep_mask("site", 4, 8, 1.0, "cpu")
    assert m.grad_checkpointing is False
    m.set_grad_checkpointing()
    assert m.grad_checkpointing is True
    m.set_grad_checkpointing(False)
    assert m.grad_checkpointing is False
