"""llm-qat_b200 — B200-native (sm_100a) implementation of LLM-QAT's
fake-quantization hot path, behind the reference's own operator API.

The directory name carries a hyphen (it mirrors the reference repo's name), so
the importable alias is ``llm_qat_b200`` (a three-line shim at the repo root).

Public surface == the reference's ``models/utils_quant.py``:
``SymQuantizer``, ``AsymQuantizer``, ``QuantizeLinear``; plus ``install()`` to
mount this implementation under the reference's import name.
"""
from __future__ import annotations

import sys

from . import _lib, fused_ops, model_patch, utils_quant
from .model_patch import fuse_kd_loss, fuse_model, mark_causal_mask, unfuse_model
from .utils_quant import AsymQuantizer, QuantizeLinear, SymQuantizer

__all__ = ["SymQuantizer", "AsymQuantizer", "QuantizeLinear", "install", "utils_quant", "fused_ops",
           "fuse_model", "unfuse_model", "fuse_kd_loss", "mark_causal_mask"]
__version__ = "0.1.0"


def install(module_name: str = "models.utils_quant") -> None:
    """Make ``from models.utils_quant import QuantizeLinear, SymQuantizer``
    (reference models/modeling_llama_quant.py:51) resolve to this package.
    Call before the reference's model file is imported."""
    sys.modules[module_name] = utils_quant
    parent, _, leaf = module_name.rpartition(".")
    if parent and parent in sys.modules:
        setattr(sys.modules[parent], leaf, utils_quant)
