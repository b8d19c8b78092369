"""Architecture registry (reference ``_get_classes``, shard/utils.py:20-30)."""
from __future__ import annotations

import importlib

from ..config import MODEL_REMAPPING, SUPPORTED_ARCHS, ModelConfig, ShardSpec  # noqa: F401
from .base import IdentityBlock, StageModel  # noqa: F401


def get_model_class(model_type: str):
    mt = MODEL_REMAPPING.get(model_type, model_type)
    if mt not in SUPPORTED_ARCHS:
        raise ValueError(f"Model type {model_type} not supported.")
    return importlib.import_module(f".{mt}", __name__).Model


def build_stage(cfg: ModelConfig, spec=None, dtype=None, device="cpu", backend=None) -> StageModel:
    import torch

    return get_model_class(cfg.model_type)(cfg, spec, dtype or torch.bfloat16, device, backend)
