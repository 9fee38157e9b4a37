"""create_model: name -> constructed module (mirrors /root/reference/models/_factory.py:46-155)."""
from __future__ import annotations

from typing import Any

import torch.nn as nn

from ._registry import is_model, model_entrypoint


def split_model_name_tag(model_name: str, no_tag: str = ""):
    model_name, *tag_list = model_name.split(".", 1)
    tag = tag_list[0] if tag_list else no_tag
    return model_name, tag


def create_model(model_name: str, pretrained: bool = False, pretrained_cfg=None, pretrained_cfg_overlay=None,
                 checkpoint_path=None, cache_dir=None, scriptable=None, exportable=None, no_jit=None,
                 **kwargs: Any) -> nn.Module:
    """Look up ``model_name``'s entrypoint and build it.

    As in the reference (``_factory.py:108``) kwargs whose value is ``None`` are dropped so that model
    defaults stay in effect.  ``pretrained=True`` raises: there is no network and no checkpoint store on
    the B200 box (the reference's hub download path is out of scope)."""
    kwargs = {k: v for k, v in kwargs.items() if v is not None}
    if ":" in model_name:
        raise NotImplementedError("hf-hub / local-dir model sources are not supported offline")
    model_name, _tag = split_model_name_tag(model_name)
    if not is_model(model_name):
        raise RuntimeError("Unknown model (%s)" % model_name)
    if pretrained:
        raise NotImplementedError("pretrained=True needs a network/checkpoint store; load a state_dict instead")
    model = model_entrypoint(model_name)(pretrained=False, **kwargs)
    if checkpoint_path:
        import torch

        sd = torch.load(checkpoint_path, map_location="cpu")
        for key in ("state_dict_ema", "model_ema", "state_dict", "model"):
            if isinstance(sd, dict) and key in sd:
                sd = sd[key]
                break
        sd = {(k[7:] if k.startswith("module.") else k): v for k, v in sd.items()}
        model.load_state_dict(sd)
    return model
