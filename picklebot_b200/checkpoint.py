"""Checkpoint compatibility: the reference saves plain ``state_dict``s
(``train.py:316-318,338``) whose keys may carry the ``_orig_mod.`` prefix of ``torch.compile`` and, when saved from
the DDP wrapper at exit, ``module.``; ``train.py:38-44`` strips the former when loading.  The modules of this
package use the reference's parameter names, so a converted checkpoint loads with ``strict=True`` in either
direction."""
from __future__ import annotations

from typing import Dict

import torch

_PREFIXES = ("_orig_mod.", "module.")


def state_dict_converter(state_dict: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
    """train.py:38-44, extended to the DDP prefix: strips leading ``_orig_mod.`` / ``module.`` (in any nesting
    order) from every key, in place, and returns the dict."""
    for key in list(state_dict.keys()):
        new_key = key
        stripped = True
        while stripped:
            stripped = False
            for pre in _PREFIXES:
                if new_key.startswith(pre):
                    new_key = new_key[len(pre):]
                    stripped = True
        if new_key != key:
            state_dict[new_key] = state_dict.pop(key)
    return state_dict


def load_reference_checkpoint(model: torch.nn.Module, path: str, map_location="cpu") -> torch.nn.Module:
    """``model.load_state_dict(state_dict_converter(torch.load(path)))`` with strict key checking
    (train.py:230-235)."""
    sd = torch.load(path, map_location=map_location, weights_only=True)
    model.load_state_dict(state_dict_converter(sd), strict=True)
    return model
