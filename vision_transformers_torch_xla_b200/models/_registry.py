"""Model registry (mirrors the call contract of /root/reference/models/_registry.py:75-136)."""
from __future__ import annotations

import fnmatch
import sys
import warnings
from typing import Any, Callable, Dict, List

_model_entrypoints: Dict[str, Callable[..., Any]] = {}
_model_to_module: Dict[str, str] = {}


def register_model(fn: Callable[..., Any]) -> Callable[..., Any]:
    mod = sys.modules[fn.__module__]
    model_name = fn.__name__
    if hasattr(mod, "__all__"):
        mod.__all__.append(model_name)
    else:
        mod.__all__ = [model_name]
    if model_name in _model_entrypoints:
        warnings.warn(f"Overwriting {model_name} in registry with {fn.__module__}.{model_name}.", stacklevel=2)
    _model_entrypoints[model_name] = fn
    _model_to_module[model_name] = fn.__module__.split(".")[-1]
    return fn


def is_model(model_name: str) -> bool:
    return model_name.split(".")[0] in _model_entrypoints


def model_entrypoint(model_name: str) -> Callable[..., Any]:
    return _model_entrypoints[model_name.split(".")[0]]


def list_models(filter: str = "") -> List[str]:
    names = sorted(_model_entrypoints)
    return fnmatch.filter(names, filter) if filter else names
