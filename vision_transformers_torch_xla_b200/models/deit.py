"""DeiT models (mirrors /root/reference/models/deit.py:28-119, 306-314) on the vitk kernels."""
from __future__ import annotations

from typing import Optional

import torch
import torch.nn as nn

from ._registry import register_model
from .vision_transformer import VisionTransformer, trunc_normal_

__all__ = ["VisionTransformerDistilled"]


class VisionTransformerDistilled(VisionTransformer):
    """Vision Transformer w/ distillation token and head (DeiT, https://arxiv.org/abs/2012.12877).

    Two prefix tokens (cls, dist), ``pos_embed`` of ``num_patches + 2``, two classifier heads.  As in the
    reference, ``forward`` returns ``(x, x_dist)`` only when ``distilled_training and training``; otherwise
    the average of the two predictions."""

    def __init__(self, *args, **kwargs):
        weight_init = kwargs.pop("weight_init", "")
        super().__init__(*args, **kwargs, weight_init="skip")
        assert self.global_pool in ("token",)
        self.num_prefix_tokens = 2
        self.dist_token = nn.Parameter(torch.zeros(1, 1, self.embed_dim))
        self.pos_embed = nn.Parameter(
            torch.zeros(1, self.patch_embed.num_patches + self.num_prefix_tokens, self.embed_dim))
        self.head_dist = nn.Linear(self.embed_dim, self.num_classes) if self.num_classes > 0 else nn.Identity()
        self.distilled_training = False  # must set this True to train w/ distillation token
        self.init_weights(weight_init)

    def init_weights(self, mode=""):
        trunc_normal_(self.dist_token, std=0.02)
        super().init_weights(mode=mode)

    @torch.jit.ignore
    def group_matcher(self, coarse=False):
        return dict(stem=r"^cls_token|pos_embed|patch_embed|dist_token",
                    blocks=[(r"^blocks\.(\d+)", None), (r"^norm", (99999,))])

    @torch.jit.ignore
    def get_classifier(self):
        return self.head, self.head_dist

    def reset_classifier(self, num_classes: int, global_pool: Optional[str] = None):
        self.num_classes = num_classes
        dev = self.cls_token.device
        self.head = nn.Linear(self.embed_dim, num_classes).to(dev) if num_classes > 0 else nn.Identity()
        self.head_dist = nn.Linear(self.embed_dim, self.num_classes).to(dev) if num_classes > 0 else nn.Identity()
        self.__dict__.pop("_vitk_store", None)

    @torch.jit.ignore
    def set_distilled_training(self, enable=True):
        self.distilled_training = enable

    def forward_head(self, x, pre_logits: bool = False) -> torch.Tensor:
        st_x, x_dist = x[:, 0], x[:, 1]
        if pre_logits:
            return (st_x + x_dist) / 2
        from ..store import get_store, use_store

        st = get_store(self)
        with use_store(st):
            a = self._linear(self.head, st_x.contiguous(), st)
            b = self._linear(self.head_dist, x_dist.contiguous(), st)
        if self.distilled_training and self.training:
            return a, b
        return (a + b) / 2

    def forward(self, x: torch.Tensor, attn_mask=None):
        out = super().forward(x, attn_mask=attn_mask)
        if isinstance(out, tuple):
            a, b = out
            if self.distilled_training and self.training and not torch.jit.is_scripting():
                # only return separate classification predictions when training in distilled mode
                return a, b
            # during standard train / finetune, inference average the classifier predictions
            return (a + b) / 2
        return out


def _create_deit(variant, pretrained=False, distilled=False, **kwargs):
    if pretrained:
        raise NotImplementedError("pretrained weights need a network; load a state_dict instead")
    for k in ("pretrained_cfg", "pretrained_cfg_overlay", "cache_dir"):
        kwargs.pop(k, None)
    model_cls = VisionTransformerDistilled if distilled else VisionTransformer
    return model_cls(**kwargs)


@register_model
def deit_tiny_patch16_224(pretrained=False, **kwargs) -> VisionTransformer:
    model_args = dict(patch_size=16, embed_dim=192, depth=12, num_heads=3)
    return _create_deit("deit_tiny_patch16_224", pretrained=pretrained, **dict(model_args, **kwargs))


@register_model
def deit_small_patch16_224(pretrained=False, **kwargs) -> VisionTransformer:
    model_args = dict(patch_size=16, embed_dim=384, depth=12, num_heads=6)
    return _create_deit("deit_small_patch16_224", pretrained=pretrained, **dict(model_args, **kwargs))


@register_model
def deit_base_patch16_224(pretrained=False, **kwargs) -> VisionTransformer:
    model_args = dict(patch_size=16, embed_dim=768, depth=12, num_heads=12)
    return _create_deit("deit_base_patch16_224", pretrained=pretrained, **dict(model_args, **kwargs))


@register_model
def deit_base_patch16_384(pretrained=False, **kwargs) -> VisionTransformer:
    model_args = dict(patch_size=16, embed_dim=768, depth=12, num_heads=12, img_size=384)
    return _create_deit("deit_base_patch16_384", pretrained=pretrained, **dict(model_args, **kwargs))


@register_model
def deit_tiny_distilled_patch16_224(pretrained=False, **kwargs) -> VisionTransformerDistilled:
    model_args = dict(patch_size=16, embed_dim=192, depth=12, num_heads=3)
    return _create_deit("deit_tiny_distilled_patch16_224", pretrained=pretrained, distilled=True,
                        **dict(model_args, **kwargs))


@register_model
def deit_small_distilled_patch16_224(pretrained=False, **kwargs) -> VisionTransformerDistilled:
    model_args = dict(patch_size=16, embed_dim=384, depth=12, num_heads=6)
    return _create_deit("deit_small_distilled_patch16_224", pretrained=pretrained, distilled=True,
                        **dict(model_args, **kwargs))


@register_model
def deit_base_distilled_patch16_224(pretrained=False, **kwargs) -> VisionTransformerDistilled:
    """DeiT-base distilled (/root/reference/models/deit.py:306-314)."""
    model_args = dict(patch_size=16, embed_dim=768, depth=12, num_heads=12)
    return _create_deit("deit_base_distilled_patch16_224", pretrained=pretrained, distilled=True,
                        **dict(model_args, **kwargs))


@register_model
def deit_base_distilled_patch16_384(pretrained=False, **kwargs) -> VisionTransformerDistilled:
    model_args = dict(patch_size=16, embed_dim=768, depth=12, num_heads=12, img_size=384)
    return _create_deit("deit_base_distilled_patch16_384", pretrained=pretrained, distilled=True,
                        **dict(model_args, **kwargs))
