"""Checkpoint / tokenizer loading shared by the two embedding classes (host side, Python as in the reference)."""
from __future__ import annotations

import json
from pathlib import Path
from typing import Dict

import torch


def load_state_dict(model_dir: str) -> Dict[str, torch.Tensor]:
    """All ``*.safetensors`` (or ``pytorch_model*.bin``) shards of a local HF checkpoint directory."""
    d = Path(model_dir)
    if not d.is_dir():
        raise FileNotFoundError(
            f"{model_dir!r} is not a local checkpoint directory (this build runs offline: download the model "
            f"first, or pass encoder=... / tokenizer=... explicitly)")
    state: Dict[str, torch.Tensor] = {}
    shards = sorted(d.glob("*.safetensors"))
    if shards:
        from safetensors.torch import load_file
        for s in shards:
            state.update(load_file(str(s)))
    else:
        for s in sorted(d.glob("pytorch_model*.bin")):
            state.update(torch.load(str(s), map_location="cpu", weights_only=True))   # tensors only: never unpickle code
    if not state:
        raise FileNotFoundError(f"no weights found under {model_dir}")
    return state


def strip_prefix(state: Dict[str, torch.Tensor], prefixes=("model.", "bert.", "0.auto_model.")) -> Dict[str, torch.Tensor]:
    out = {}
    for k, v in state.items():
        for p in prefixes:
            if k.startswith(p):
                k = k[len(p):]
                break
        out[k] = v
    return out


def load_config(model_dir: str) -> dict:
    return json.loads((Path(model_dir) / "config.json").read_text())


def load_tokenizer(model_dir: str):
    from transformers import AutoTokenizer
    return AutoTokenizer.from_pretrained(model_dir, trust_remote_code=True)
