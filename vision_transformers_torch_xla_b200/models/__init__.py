"""``from models import create_model`` — same import contract as /root/reference/models/__init__.py."""
from . import deit, my_vit, vision_transformer  # noqa: F401  (registers the entrypoints)
from ._factory import create_model
from ._registry import is_model, list_models, model_entrypoint, register_model
from .deit import VisionTransformerDistilled
from .vision_transformer import (Attention, Block, DropPath, LayerNorm, LayerScale, Mlp, PatchEmbed,
                                 VisionTransformer, global_pool_nlc)

__all__ = ["create_model", "register_model", "is_model", "list_models", "model_entrypoint", "VisionTransformer",
           "VisionTransformerDistilled", "Block", "Attention", "Mlp", "PatchEmbed", "LayerNorm", "LayerScale",
           "DropPath", "global_pool_nlc"]
