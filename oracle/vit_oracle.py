"""CPU oracle for the ViT training hot path — TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A plain eager-PyTorch fp32 restatement of the reference's algorithm
(TaiMingLu/vision_transformers_torch_xla).  Only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import this package; the product
(``vision_transformers_torch_xla_b200``) never does and has no CPU fallback.

PARITY PIN.  The reference cannot be imported as a package offline (timm / torch_xla / tensorflow are
absent), and its own tests (test_kd.py) assert no numerical values.  It is pinned in two halves:

* IN-TREE HALF — PINNED TO THE REFERENCE'S OWN CODE: everything the reference itself defines on the hot
  path (LayerScale, Block, global_pool_nlc, VisionTransformer ctor wiring / init order / _pos_embed / forward,
  the vit_* / my_vit_* / deit_* entrypoints, VisionTransformerDistilled, cosine_scheduler,
  get_parameter_groups / create_optimizer, the closure-local DistillationLoss / StudentWithDistillation,
  engine.train_one_epoch's eager branch with its schedule write, engine.evaluate) is ast-extracted from
  /root/reference and EXECUTED by tests/golden/make_ref_fixtures.py; tests/test_ref_fixtures.py holds this
  file to those outputs (logits, per-block activations, every gradient, seeded-init checksums, schedules,
  group membership, loss values, 10-step engine runs) at <= 1e-6.
* LEAF HALF — UNPINNED: the leaf arithmetic (Attention, Mlp, PatchEmbed, LayerNorm, DropPath,
  SoftTargetCrossEntropy, Mixup, accuracy) lives in the un-vendored pip package ``timm==1.0.15``
  (/root/reference/requirements.txt:13, resolved through /root/reference/models/_compat.py:27-172) whose
  source is not in the container.  These few classes restate timm 1.0.15's published semantics (SURVEY.md
  Appendix A.2) and are cross-checked against two independent implementations of the same architecture
  (``torchvision.models.VisionTransformer`` and ``transformers.ViTForImageClassification`` on shared
  weights, tests/test_oracle.py).

Each class/function cites the reference lines it follows.
"""
from __future__ import annotations

import math
from typing import Dict, Iterable, List, Set, Tuple

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F


# --------------------------------------------------------------------------------------------
# leaf layers (timm 1.0.15 semantics)
# --------------------------------------------------------------------------------------------
def trunc_normal_(tensor: torch.Tensor, mean: float = 0.0, std: float = 1.0, a: float = -2.0, b: float = 2.0):
    """timm.layers.trunc_normal_: truncation bounds are ABSOLUTE (+-2), i.e. +-100 sigma at std=.02."""
    return nn.init.trunc_normal_(tensor, mean=mean, std=std, a=a, b=b)


class LayerNorm(nn.LayerNorm):
    """timm.layers.LayerNorm == nn.LayerNorm with eps=1e-6 (used at vision_transformer.py:148,163,603,616)."""

    def __init__(self, num_channels: int, eps: float = 1e-6, affine: bool = True):
        super().__init__(num_channels, eps=eps, elementwise_affine=affine)


class DropPath(nn.Module):
    """timm.layers.DropPath (per-sample stochastic depth, scale_by_keep=True)."""

    def __init__(self, drop_prob: float = 0.0, scale_by_keep: bool = True):
        super().__init__()
        self.drop_prob = drop_prob
        self.scale_by_keep = scale_by_keep

    def forward(self, x):
        if self.drop_prob == 0.0 or not self.training:
            return x
        keep_prob = 1 - self.drop_prob
        shape = (x.shape[0],) + (1,) * (x.ndim - 1)
        random_tensor = x.new_empty(shape).bernoulli_(keep_prob)
        if keep_prob > 0.0 and self.scale_by_keep:
            random_tensor.div_(keep_prob)
        return x * random_tensor


class PatchEmbed(nn.Module):
    """timm.layers.PatchEmbed: Conv2d(k=s=patch) -> flatten(2).transpose(1,2); ctor call vision_transformer.py:552-560."""

    def __init__(self, img_size=224, patch_size=16, in_chans=3, embed_dim=768, bias=True):
        super().__init__()
        self.img_size = (img_size, img_size) if isinstance(img_size, int) else tuple(img_size)
        self.patch_size = (patch_size, patch_size) if isinstance(patch_size, int) else tuple(patch_size)
        self.grid_size = (self.img_size[0] // self.patch_size[0], self.img_size[1] // self.patch_size[1])
        self.num_patches = self.grid_size[0] * self.grid_size[1]
        self.proj = nn.Conv2d(in_chans, embed_dim, kernel_size=self.patch_size, stride=self.patch_size, bias=bias)
        self.norm = nn.Identity()

    def forward(self, x):
        B, C, H, W = x.shape
        assert H == self.img_size[0] and W == self.img_size[1], "Input size doesn't match model"
        x = self.proj(x)
        x = x.flatten(2).transpose(1, 2)
        return self.norm(x)


class Attention(nn.Module):
    """timm Attention (witness of the same code shape in-tree: /root/reference/models/eva.py:146-194)."""

    def __init__(self, dim, num_heads=8, qkv_bias=False, proj_bias=True, attn_drop=0.0, proj_drop=0.0, fused=True):
        super().__init__()
        assert dim % num_heads == 0
        self.num_heads = num_heads
        self.head_dim = dim // num_heads
        self.scale = self.head_dim ** -0.5
        self.fused_attn = fused
        self.qkv = nn.Linear(dim, dim * 3, bias=qkv_bias)
        self.q_norm = nn.Identity()
        self.k_norm = nn.Identity()
        self.attn_drop = nn.Dropout(attn_drop)
        self.proj = nn.Linear(dim, dim, bias=proj_bias)
        self.proj_drop = nn.Dropout(proj_drop)

    def forward(self, x, attn_mask=None):
        B, N, C = x.shape
        qkv = self.qkv(x).reshape(B, N, 3, self.num_heads, self.head_dim).permute(2, 0, 3, 1, 4)
        q, k, v = qkv.unbind(0)
        q, k = self.q_norm(q), self.k_norm(k)
        if self.fused_attn:
            x = F.scaled_dot_product_attention(q, k, v, attn_mask=attn_mask,
                                               dropout_p=self.attn_drop.p if self.training else 0.0)
        else:
            q = q * self.scale
            attn = q @ k.transpose(-2, -1)
            attn = attn.softmax(dim=-1)
            attn = self.attn_drop(attn)
            x = attn @ v
        x = x.transpose(1, 2).reshape(B, N, C)
        x = self.proj(x)
        return self.proj_drop(x)


class Mlp(nn.Module):
    """timm.layers.Mlp: fc1 -> GELU(erf) -> drop -> fc2 -> drop; ctor call vision_transformer.py:164-171."""

    def __init__(self, in_features, hidden_features=None, out_features=None, bias=True, drop=0.0):
        super().__init__()
        out_features = out_features or in_features
        hidden_features = hidden_features or in_features
        self.fc1 = nn.Linear(in_features, hidden_features, bias=bias)
        self.act = nn.GELU()
        self.drop1 = nn.Dropout(drop)
        self.norm = nn.Identity()
        self.fc2 = nn.Linear(hidden_features, out_features, bias=bias)
        self.drop2 = nn.Dropout(drop)

    def forward(self, x):
        return self.drop2(self.fc2(self.norm(self.drop1(self.act(self.fc1(x))))))


# --------------------------------------------------------------------------------------------
# in-tree model code (cite-able)
# --------------------------------------------------------------------------------------------
class LayerScale(nn.Module):
    """/root/reference/models/vision_transformer.py:80-106."""

    def __init__(self, dim: int, init_values: float = 1e-5):
        super().__init__()
        self.gamma = nn.Parameter(init_values * torch.ones(dim))

    def forward(self, x):
        return x * self.gamma


class Block(nn.Module):
    """/root/reference/models/vision_transformer.py:109-178 (pre-norm; residual after DropPath(LayerScale(branch)))."""

    def __init__(self, dim, num_heads, mlp_ratio=4.0, qkv_bias=False, proj_bias=True, proj_drop=0.0, attn_drop=0.0,
                 init_values=None, drop_path=0.0, fused_attn=True):
        super().__init__()
        self.norm1 = LayerNorm(dim)
        self.attn = Attention(dim, num_heads=num_heads, qkv_bias=qkv_bias, proj_bias=proj_bias, attn_drop=attn_drop,
                              proj_drop=proj_drop, fused=fused_attn)
        self.ls1 = LayerScale(dim, init_values=init_values) if init_values else nn.Identity()
        self.drop_path1 = DropPath(drop_path) if drop_path > 0.0 else nn.Identity()
        self.norm2 = LayerNorm(dim)
        self.mlp = Mlp(in_features=dim, hidden_features=int(dim * mlp_ratio), bias=proj_bias, drop=proj_drop)
        self.ls2 = LayerScale(dim, init_values=init_values) if init_values else nn.Identity()
        self.drop_path2 = DropPath(drop_path) if drop_path > 0.0 else nn.Identity()

    def forward(self, x, attn_mask=None):
        x = x + self.drop_path1(self.ls1(self.attn(self.norm1(x), attn_mask=attn_mask)))
        x = x + self.drop_path2(self.ls2(self.mlp(self.norm2(x))))
        return x


def global_pool_nlc(x, pool_type="token", num_prefix_tokens=1):
    """/root/reference/models/vision_transformer.py:419-441 ('token' and 'avg' only)."""
    if not pool_type:
        return x
    if pool_type == "token":
        return x[:, 0]
    assert pool_type == "avg", pool_type
    return x[:, num_prefix_tokens:].mean(dim=1)


def init_weights_vit_timm(module: nn.Module) -> None:
    """/root/reference/models/vision_transformer.py:998-1010 (only nn.Linear is touched)."""
    if isinstance(module, nn.Linear):
        trunc_normal_(module.weight, std=0.02)
        if module.bias is not None:
            nn.init.zeros_(module.bias)


class VisionTransformer(nn.Module):
    """/root/reference/models/vision_transformer.py:444-995 restricted to the feature set of the five configs."""

    def __init__(self, img_size=224, patch_size=16, in_chans=3, num_classes=1000, global_pool="token", embed_dim=768,
                 depth=12, num_heads=12, mlp_ratio=4.0, qkv_bias=True, proj_bias=True, init_values=None,
                 class_token=True, final_norm=True, fc_norm=None, drop_rate=0.0, pos_drop_rate=0.0,
                 proj_drop_rate=0.0, attn_drop_rate=0.0, drop_path_rate=0.0, weight_init="", fused_attn=True):
        super().__init__()
        assert global_pool in ("", "avg", "token")
        assert class_token or global_pool != "token"
        use_fc_norm = global_pool in ("avg",) if fc_norm is None else fc_norm  # :529
        self.num_classes = num_classes
        self.global_pool = global_pool
        self.num_features = self.head_hidden_size = self.embed_dim = embed_dim
        self.num_prefix_tokens = 1 if class_token else 0  # :537
        self.has_class_token = class_token
        self.patch_embed = PatchEmbed(img_size=img_size, patch_size=patch_size, in_chans=in_chans, embed_dim=embed_dim,
                                      bias=True)  # :552-560 (bias = not pre_norm)
        num_patches = self.patch_embed.num_patches
        self.cls_token = nn.Parameter(torch.zeros(1, 1, embed_dim)) if class_token else None  # :564
        embed_len = num_patches + self.num_prefix_tokens
        self.pos_embed = nn.Parameter(torch.randn(1, embed_len, embed_dim) * 0.02)  # :566-570
        self.pos_drop = nn.Dropout(p=pos_drop_rate)
        dpr = [x.item() for x in torch.linspace(0, drop_path_rate, depth, device="cpu")]  # :581
        self.blocks = nn.Sequential(*[
            Block(dim=embed_dim, num_heads=num_heads, mlp_ratio=mlp_ratio, qkv_bias=qkv_bias, proj_bias=proj_bias,
                  init_values=init_values, proj_drop=proj_drop_rate, attn_drop=attn_drop_rate, drop_path=dpr[i],
                  fused_attn=fused_attn)
            for i in range(depth)])
        self.norm = LayerNorm(embed_dim) if final_norm and not use_fc_norm else nn.Identity()  # :603
        self.fc_norm = LayerNorm(embed_dim) if final_norm and use_fc_norm else nn.Identity()  # :616
        self.head_drop = nn.Dropout(drop_rate)
        self.head = nn.Linear(self.embed_dim, num_classes) if num_classes > 0 else nn.Identity()  # :618
        if weight_init != "skip":
            self.init_weights(weight_init)

    def init_weights(self, mode: str = "") -> None:  # :634-648
        assert mode == ""
        if self.pos_embed is not None:
            trunc_normal_(self.pos_embed, std=0.02)
        if self.cls_token is not None:
            nn.init.normal_(self.cls_token, std=1e-6)
        self.apply(init_weights_vit_timm)

    def no_weight_decay(self) -> Set[str]:  # :665-668
        return {"pos_embed", "cls_token", "dist_token"}

    def get_classifier(self):
        return self.head

    def _pos_embed(self, x):  # :743-780 (no_embed_class=False branch)
        to_cat = []
        if self.cls_token is not None:
            to_cat.append(self.cls_token.expand(x.shape[0], -1, -1))
        if to_cat:
            x = torch.cat(to_cat + [x], dim=1)
        x = x + self.pos_embed
        return self.pos_drop(x)

    def forward_features(self, x):  # :934-951
        x = self.patch_embed(x)
        x = self._pos_embed(x)
        x = self.blocks(x)
        return self.norm(x)

    def forward_head(self, x, pre_logits: bool = False):  # :977-990
        x = global_pool_nlc(x, pool_type=self.global_pool, num_prefix_tokens=self.num_prefix_tokens)
        x = self.fc_norm(x)
        x = self.head_drop(x)
        return x if pre_logits else self.head(x)

    def forward(self, x):  # :992-995
        return self.forward_head(self.forward_features(x))


class VisionTransformerDistilled(VisionTransformer):
    """/root/reference/models/deit.py:28-119."""

    def __init__(self, *args, **kwargs):
        weight_init = kwargs.pop("weight_init", "")
        super().__init__(*args, **kwargs, weight_init="skip")
        assert self.global_pool in ("token",)
        self.num_prefix_tokens = 2
        self.dist_token = nn.Parameter(torch.zeros(1, 1, self.embed_dim))
        self.pos_embed = nn.Parameter(torch.zeros(1, self.patch_embed.num_patches + self.num_prefix_tokens, self.embed_dim))
        self.head_dist = nn.Linear(self.embed_dim, self.num_classes) if self.num_classes > 0 else nn.Identity()
        self.distilled_training = False
        self.init_weights(weight_init)

    def init_weights(self, mode=""):
        trunc_normal_(self.dist_token, std=0.02)
        super().init_weights(mode=mode)

    def set_distilled_training(self, enable=True):
        self.distilled_training = enable

    def get_classifier(self):
        return self.head, self.head_dist

    def _pos_embed(self, x):  # deit.py:75-106
        x = torch.cat((self.cls_token.expand(x.shape[0], -1, -1), self.dist_token.expand(x.shape[0], -1, -1), x), dim=1)
        x = x + self.pos_embed
        return self.pos_drop(x)

    def forward_head(self, x, pre_logits: bool = False):  # deit.py:108-119
        x, x_dist = x[:, 0], x[:, 1]
        if pre_logits:
            return (x + x_dist) / 2
        x = self.head(x)
        x_dist = self.head_dist(x_dist)
        if self.distilled_training and self.training:
            return x, x_dist
        return (x + x_dist) / 2


#: model sizes: /root/reference/models/vision_transformer.py:2690-2860 (vit_*), my_vit.py:84-165, deit.py:306-314
MODEL_CFGS: Dict[str, dict] = {
    "vit_tiny_patch16_224": dict(patch_size=16, embed_dim=192, depth=12, num_heads=3),
    "vit_small_patch16_224": dict(patch_size=16, embed_dim=384, depth=12, num_heads=6),
    "vit_base_patch16_224": dict(patch_size=16, embed_dim=768, depth=12, num_heads=12),
    "vit_large_patch16_224": dict(patch_size=16, embed_dim=1024, depth=24, num_heads=16),
    "vit_large_patch16_384": dict(patch_size=16, embed_dim=1024, depth=24, num_heads=16, img_size=384),
    "my_vit_mini": dict(patch_size=16, embed_dim=144, depth=12, num_heads=3),
    "my_vit_ti": dict(patch_size=16, embed_dim=192, depth=12, num_heads=3),
    "my_vit_xs": dict(patch_size=16, embed_dim=288, depth=12, num_heads=4),
    "my_vit_s": dict(patch_size=16, embed_dim=384, depth=12, num_heads=6),
    "my_vit_b": dict(patch_size=16, embed_dim=768, depth=12, num_heads=12),
    "my_vit_l": dict(patch_size=16, embed_dim=1024, depth=24, num_heads=16),
    "deit_base_distilled_patch16_224": dict(patch_size=16, embed_dim=768, depth=12, num_heads=12, distilled=True),
    "deit_tiny_distilled_patch16_224": dict(patch_size=16, embed_dim=192, depth=12, num_heads=3, distilled=True),
}


def create_model(model_name: str, pretrained: bool = False, **kwargs) -> nn.Module:
    """/root/reference/models/_factory.py:46-155 reduced to name -> ctor kwargs (None kwargs pruned, :108)."""
    assert not pretrained, "no network/checkpoints offline"
    kwargs = {k: v for k, v in kwargs.items() if v is not None}
    cfg = dict(MODEL_CFGS[model_name])
    distilled = cfg.pop("distilled", False)
    cfg.update(kwargs)
    cls = VisionTransformerDistilled if distilled else VisionTransformer
    return cls(**cfg)


# --------------------------------------------------------------------------------------------
# losses (timm.loss + /root/reference/main.py:836-850, 926-968)
# --------------------------------------------------------------------------------------------
class SoftTargetCrossEntropy(nn.Module):
    def forward(self, x, target):
        loss = torch.sum(-target * F.log_softmax(x, dim=-1), dim=-1)
        return loss.mean()


class LabelSmoothingCrossEntropy(nn.Module):
    def __init__(self, smoothing=0.1):
        super().__init__()
        assert smoothing < 1.0
        self.smoothing = smoothing
        self.confidence = 1.0 - smoothing

    def forward(self, x, target):
        logprobs = F.log_softmax(x, dim=-1)
        nll_loss = -logprobs.gather(dim=-1, index=target.unsqueeze(1)).squeeze(1)
        smooth_loss = -logprobs.mean(dim=-1)
        loss = self.confidence * nll_loss + self.smoothing * smooth_loss
        return loss.mean()


class DistillationLoss(nn.Module):
    """/root/reference/main.py:939-968 (soft KD); ``hard=True`` is the DeiT-paper hard variant."""

    def __init__(self, base_criterion, alpha=0.7, temperature=4.0, hard=False):
        super().__init__()
        self.base_criterion = base_criterion
        self.alpha = alpha
        self.temperature = temperature
        self.hard = hard
        self.kl_div = nn.KLDivLoss(reduction="batchmean")

    def forward(self, outputs, targets):
        if isinstance(outputs, tuple):
            student_logits, teacher_logits = outputs
            dist_logits = student_logits
            if isinstance(student_logits, tuple):  # DeiT distilled student: (cls_logits, dist_logits)
                student_logits, dist_logits = student_logits
            ce_loss = self.base_criterion(student_logits, targets)
            if self.hard:
                kd_loss = F.cross_entropy(dist_logits, teacher_logits.argmax(dim=1))
            else:
                student_soft = torch.log_softmax(dist_logits / self.temperature, dim=1)
                teacher_soft = torch.softmax(teacher_logits / self.temperature, dim=1)
                kd_loss = self.kl_div(student_soft, teacher_soft) * (self.temperature ** 2)
            return (1 - self.alpha) * ce_loss + self.alpha * kd_loss
        return self.base_criterion(outputs, targets)


class StudentWithDistillation(nn.Module):
    """/root/reference/main.py:836-850."""

    def __init__(self, student_model, teacher_model):
        super().__init__()
        self.student = student_model
        self.teacher = teacher_model

    def forward(self, x):
        student_logits = self.student(x)
        if self.training and self.teacher is not None:
            with torch.no_grad():
                teacher_logits = self.teacher(x)
            return student_logits, teacher_logits
        return student_logits


# --------------------------------------------------------------------------------------------
# optimizer groups, schedules, train step (optim_factory.py, utils/__init__.py, engine.py)
# --------------------------------------------------------------------------------------------
def get_parameter_groups(model: nn.Module, weight_decay: float = 1e-5, skip_list: Iterable[str] = (),
                         tpu_name_rule: bool = False) -> List[dict]:
    """/root/reference/optim_factory.py:155-195 (non-TPU shape rule), lr_scale = 1.  ``tpu_name_rule`` selects the
    name-only rule the reference uses when ``PJRT_DEVICE=TPU`` (:104-106): the two differ
    only for 1-D parameters whose name carries neither ``.bias`` nor ``norm``/``bn`` (LayerScale ``gamma``)."""
    groups: Dict[str, dict] = {}
    for name, param in model.named_parameters():
        if not param.requires_grad:
            continue
        if tpu_name_rule:
            no_decay = (name.endswith(".bias") or name.endswith(".weight") and ("norm" in name.lower() or "bn" in name.lower())
                        or name in skip_list)
        else:
            no_decay = len(param.shape) == 1 or name.endswith(".bias") or name in skip_list
        if no_decay:
            group_name, this_wd = "no_decay", 0.0
        else:
            group_name, this_wd = "decay", weight_decay
        if group_name not in groups:
            groups[group_name] = {"weight_decay": this_wd, "params": [], "lr_scale": 1.0}
        groups[group_name]["params"].append(param)
    return list(groups.values())


def create_optimizer(model: nn.Module, lr: float, weight_decay: float, eps: float = 1e-8,
                     betas: Tuple[float, float] = (0.9, 0.999)) -> torch.optim.Optimizer:
    """/root/reference/optim_factory.py:214-249 for --opt adamw (filter_bias_and_bn=True)."""
    skip = model.no_weight_decay() if hasattr(model, "no_weight_decay") else {}
    parameters = get_parameter_groups(model, weight_decay, skip)
    return torch.optim.AdamW(parameters, lr=lr, weight_decay=0.0, eps=eps, betas=betas)


def cosine_scheduler(base_value, final_value, epochs, niter_per_ep, warmup_epochs=0, start_warmup_value=0,
                     warmup_steps=-1) -> np.ndarray:
    """/root/reference/utils/__init__.py:667-684."""
    warmup_schedule = np.array([])
    warmup_iters = warmup_epochs * niter_per_ep
    if warmup_steps > 0:
        warmup_iters = warmup_steps
    if warmup_epochs > 0:
        warmup_schedule = np.linspace(start_warmup_value, base_value, warmup_iters)
    iters = np.arange(epochs * niter_per_ep - warmup_iters)
    schedule = np.array(
        [final_value + 0.5 * (base_value - final_value) * (1 + math.cos(math.pi * i / (len(iters)))) for i in iters])
    schedule = np.concatenate((warmup_schedule, schedule))
    assert len(schedule) == epochs * niter_per_ep
    return schedule


def apply_schedules(optimizer, it: int, lr_schedule_values=None, wd_schedule_values=None, wd_quirk: bool = True) -> None:
    """/root/reference/engine.py:98-103.  ``wd_quirk=True`` reproduces the reference as written: the test
    ``param_group.get("weight_decay") is not None`` is also true for 0.0, so the no_decay group is overwritten
    (SURVEY Appendix D #1); ``False`` is the commented-out original (``> 0``, engine.py:466)."""
    for param_group in optimizer.param_groups:
        if lr_schedule_values is not None:
            param_group["lr"] = lr_schedule_values[it] * param_group.get("lr_scale", 1.0)
        if wd_schedule_values is not None:
            wd = param_group.get("weight_decay", None)
            if (wd is not None) if wd_quirk else (wd is not None and wd > 0):
                param_group["weight_decay"] = wd_schedule_values[it]


def train_step(model, criterion, optimizer, samples, targets, update_freq: int = 1, do_step: bool = True):
    """/root/reference/engine.py:257-274 (eager fp32 branch): fwd, loss/update_freq, backward, step, zero_grad."""
    output = model(samples)
    loss = criterion(output, targets)
    loss = loss / update_freq
    loss.backward()
    if do_step:
        optimizer.step()
        optimizer.zero_grad()
    return loss.detach(), output


def train_one_epoch(model, criterion, data_loader, optimizer, epoch: int = 0, start_steps: int = 0,
                    lr_schedule_values=None, wd_schedule_values=None, num_training_steps_per_epoch=None,
                    update_freq: int = 1, wd_quirk: bool = True) -> Dict[str, float]:
    """Eager branch of /root/reference/engine.py:19-333 without logging/EMA/mixup (host-side, out of scope)."""
    model.train(True)
    optimizer.zero_grad()
    losses, lrs = [], []
    for data_iter_step, (samples, targets) in enumerate(data_loader):
        step = data_iter_step // update_freq
        if num_training_steps_per_epoch is not None and step >= num_training_steps_per_epoch:
            continue
        it = start_steps + step
        if (lr_schedule_values is not None or wd_schedule_values is not None) and data_iter_step % update_freq == 0:
            apply_schedules(optimizer, it, lr_schedule_values, wd_schedule_values, wd_quirk)
        loss, _ = train_step(model, criterion, optimizer, samples, targets, update_freq,
                             do_step=(data_iter_step + 1) % update_freq == 0)
        losses.append(float(loss))
        lrs.append(optimizer.param_groups[0]["lr"])   # engine.py:299-300: the lr meter is fed every logged iteration
    return {"loss": float(np.mean(losses)) if losses else float("nan"), "lr": float(np.mean(lrs)) if lrs else float("nan"),
            "losses": losses}


def accuracy(output: torch.Tensor, target: torch.Tensor, topk=(1,)):
    """timm.utils.accuracy (1.0.15): top-k precision in percent."""
    maxk = min(max(topk), output.size()[1])
    batch_size = target.size(0)
    _, pred = output.topk(maxk, 1, True, True)
    pred = pred.t()
    correct = pred.eq(target.reshape(1, -1).expand_as(pred))
    return [correct[:min(k, maxk)].reshape(-1).float().sum(0) * 100.0 / batch_size for k in topk]


@torch.no_grad()
def evaluate(data_loader, model) -> Dict[str, float]:
    """/root/reference/engine.py:339-430 without the TPU plumbing: CrossEntropyLoss + top-1 / top-5 meters; the loss
    meter averages per batch (``update(loss=...)`` with n=1), the accuracy meters per sample (n=batch_size)."""
    criterion = nn.CrossEntropyLoss()
    model.eval()
    tot = {"loss": [0.0, 0], "acc1": [0.0, 0], "acc5": [0.0, 0]}
    for batch in data_loader:
        images, target = batch[0], batch[-1]
        output = model(images)
        loss = criterion(output, target)
        acc1, acc5 = accuracy(output, target, topk=(1, 5))
        n = images.shape[0]
        for k, v, w in (("loss", loss.item(), 1), ("acc1", acc1.item(), n), ("acc5", acc5.item(), n)):
            tot[k][0] += v * w
            tot[k][1] += w
    return {k: v[0] / v[1] for k, v in tot.items()}


def mixup_soft_targets(labels: torch.Tensor, num_classes: int = 1000, lam: float = 0.7, smoothing: float = 0.1):
    """timm.data.Mixup target transform: lam*onehot_s(y) + (1-lam)*onehot_s(y.flip(0))."""
    off = smoothing / num_classes
    on = 1.0 - smoothing + off

    def one_hot(y):
        return torch.full((y.shape[0], num_classes), off).scatter_(1, y.view(-1, 1), on)

    return lam * one_hot(labels) + (1.0 - lam) * one_hot(labels.flip(0))


class Mixup:
    """Restatement of ``timm.data.Mixup`` 1.0.15, batch mode, as the reference constructs and calls it
    (/root/reference/main.py:622-629; engine.py:259-262).  Pure torch on whatever device ``x`` lives on; draws
    ``np.random`` numbers in timm's order: rand() for prob, [rand() for the cutmix switch,] beta() for lam,
    [randint() twice for the box centre]."""

    def __init__(self, mixup_alpha=1.0, cutmix_alpha=0.0, cutmix_minmax=None, prob=1.0, switch_prob=0.5, mode="batch",
                 correct_lam=True, label_smoothing=0.1, num_classes=1000):
        assert cutmix_minmax is None and mode == "batch"
        self.mixup_alpha, self.cutmix_alpha, self.mix_prob, self.switch_prob = mixup_alpha, cutmix_alpha, prob, switch_prob
        self.label_smoothing, self.num_classes, self.correct_lam = label_smoothing, num_classes, correct_lam

    def _params_per_batch(self):
        import numpy as np
        lam, use_cutmix = 1.0, False
        if np.random.rand() < self.mix_prob:
            if self.mixup_alpha > 0.0 and self.cutmix_alpha > 0.0:
                use_cutmix = np.random.rand() < self.switch_prob
                lam_mix = np.random.beta(self.cutmix_alpha, self.cutmix_alpha) if use_cutmix else \
                    np.random.beta(self.mixup_alpha, self.mixup_alpha)
            elif self.mixup_alpha > 0.0:
                lam_mix = np.random.beta(self.mixup_alpha, self.mixup_alpha)
            else:
                use_cutmix = True
                lam_mix = np.random.beta(self.cutmix_alpha, self.cutmix_alpha)
            lam = float(lam_mix)
        return lam, use_cutmix

    def __call__(self, x, target):
        import numpy as np
        assert len(x) % 2 == 0
        lam, use_cutmix = self._params_per_batch()
        if lam != 1.0:
            if use_cutmix:
                img_h, img_w = x.shape[-2:]
                ratio = np.sqrt(1 - lam)
                cut_h, cut_w = int(img_h * ratio), int(img_w * ratio)
                cy = np.random.randint(0, img_h)
                cx = np.random.randint(0, img_w)
                yl, yh = np.clip(cy - cut_h // 2, 0, img_h), np.clip(cy + cut_h // 2, 0, img_h)
                xl, xh = np.clip(cx - cut_w // 2, 0, img_w), np.clip(cx + cut_w // 2, 0, img_w)
                if self.correct_lam:
                    lam = 1.0 - (yh - yl) * (xh - xl) / float(img_h * img_w)
                x[:, :, yl:yh, xl:xh] = x.flip(0)[:, :, yl:yh, xl:xh]
            else:
                x_flipped = x.flip(0).mul_(1.0 - lam)
                x.mul_(lam).add_(x_flipped)
        off = self.label_smoothing / self.num_classes
        on = 1.0 - self.label_smoothing + off
        t = target.long().view(-1, 1)
        y1 = torch.full((t.shape[0], self.num_classes), off, device=x.device).scatter_(1, t.to(x.device), on)
        y2 = torch.full((t.shape[0], self.num_classes), off, device=x.device).scatter_(1, t.flip(0).to(x.device), on)
        return x, y1 * lam + y2 * (1.0 - lam)
