"""Custom ViT sizes trained by the reference's launch scripts (mirrors /root/reference/models/my_vit.py:84-165;
``my_vit_b`` is the model of run_train.sh:56)."""
from __future__ import annotations

from ._registry import register_model
from .vision_transformer import VisionTransformer, _create_vision_transformer


@register_model
def my_vit_mini(pretrained: bool = False, **kwargs) -> VisionTransformer:
    """ViT-Mini/16 — ~3.3 M params (head_dim 48: runs on the 64-wide attention tiles, zero-padded by TMA)."""
    model_args = dict(patch_size=16, embed_dim=144, depth=12, num_heads=3)
    return _create_vision_transformer("my_vit_mini", pretrained=pretrained, **dict(model_args, **kwargs))


@register_model
def my_vit_ti(pretrained: bool = False, **kwargs) -> VisionTransformer:
    """ViT-Tiny/16 — 5.7 M params"""
    model_args = dict(patch_size=16, embed_dim=192, depth=12, num_heads=3)
    return _create_vision_transformer("my_vit_ti", pretrained=pretrained, **dict(model_args, **kwargs))


@register_model
def my_vit_xs(pretrained: bool = False, **kwargs) -> VisionTransformer:
    """ViT-XS/16 — ~11 M params (head_dim 72: every attention operand takes a second, zero-padded tile; N <= 256)."""
    model_args = dict(patch_size=16, embed_dim=288, depth=12, num_heads=4)
    return _create_vision_transformer("my_vit_xs", pretrained=pretrained, **dict(model_args, **kwargs))


@register_model
def my_vit_s(pretrained: bool = False, **kwargs) -> VisionTransformer:
    """ViT-Small/16 — 22 M params"""
    model_args = dict(patch_size=16, embed_dim=384, depth=12, num_heads=6)
    return _create_vision_transformer("my_vit_s", pretrained=pretrained, **dict(model_args, **kwargs))


@register_model
def my_vit_b(pretrained: bool = False, **kwargs) -> VisionTransformer:
    """ViT-Base/16 — 86 M params"""
    model_args = dict(patch_size=16, embed_dim=768, depth=12, num_heads=12)
    return _create_vision_transformer("my_vit_b", pretrained=pretrained, **dict(model_args, **kwargs))


@register_model
def my_vit_l(pretrained: bool = False, **kwargs) -> VisionTransformer:
    """ViT-Large/16 — 304 M params"""
    model_args = dict(patch_size=16, embed_dim=1024, depth=24, num_heads=16)
    return _create_vision_transformer("my_vit_l", pretrained=pretrained, **dict(model_args, **kwargs))
