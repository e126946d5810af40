"""CPU: pin the encoder oracle against vectors produced by the reference's own vendored Qwen2Model."""
from pathlib import Path

import numpy as np
import pytest
import torch

from oracle import encoder as oenc
from easyrag_b200.encoder import Qwen2Config, BertConfig, random_state

GOLD = Path(__file__).parent / "golden" / "qwen2_tiny.npz"


def load_golden():
    z = np.load(GOLD)
    c = z["cfg"]
    cfg = Qwen2Config(vocab_size=int(c[0]), hidden_size=int(c[1]), intermediate_size=int(c[2]),
                      num_hidden_layers=int(c[3]), num_attention_heads=int(c[4]), num_key_value_heads=int(c[5]),
                      max_position_embeddings=int(c[6]), rms_norm_eps=float(z["rms_norm_eps"][0]),
                      rope_theta=float(z["rope_theta"][0]))
    state = {k[3:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("w::")}
    return z, cfg, state


def test_oracle_reproduces_reference_model_fp32():
    z, cfg, state = load_golden()
    ids, mask = torch.from_numpy(z["input_ids"]), torch.from_numpy(z["attention_mask"])
    got = oenc.gte_embed(state, cfg, ids, mask, torch.float32).numpy()
    assert np.abs(got - z["emb_fp32"]).max() < 2e-6           # same math, same weights: float32 round-off only


def test_oracle_bf16_tracks_reference_bf16():
    z, cfg, state = load_golden()
    ids, mask = torch.from_numpy(z["input_ids"]), torch.from_numpy(z["attention_mask"])
    got = oenc.gte_embed(state, cfg, ids, mask, torch.bfloat16).numpy()
    cos = (got * z["emb_bf16"]).sum(1) / np.linalg.norm(got, axis=1) / np.linalg.norm(z["emb_bf16"], axis=1)
    assert cos.min() > 1 - 1e-3


def test_padding_side_does_not_change_embeddings():
    # RoPE is relative: left-padded (column positions) and right-padded batches give the same vectors
    z, cfg, state = load_golden()
    seqs = [[int(t) for t, m in zip(r, mk) if m] for r, mk in zip(z["input_ids"], z["attention_mask"])]
    a = oenc.gte_embed(state, cfg, *oenc.pad_left(seqs))
    b = oenc.gte_embed(state, cfg, *oenc.pad_right(seqs))
    assert (a - b).abs().max() < 1e-4


def test_bert_oracle_runs():
    cfg = BertConfig(vocab_size=300, hidden_size=128, intermediate_size=256, num_hidden_layers=2,
                     num_attention_heads=2, max_position_embeddings=64)
    st = random_state("bert", cfg, 1, std=0.05)
    e = oenc.bert_embed(st, cfg, [[5, 6, 7], [9] * 20], pooling="cls")
    assert e.shape == (2, 128) and np.allclose(e.norm(dim=1).numpy(), 1.0, atol=1e-5)
