#!/usr/bin/env python
"""Bandwidth of the encoder's element-wise kernels (LayerNorm, RMSNorm, RoPE) at the token counts of the encode bench.

    python scripts/bench_ops.py [--rows 147456]       # one JSON line per kernel: ms, algorithmic GB/s
"""
import argparse
import json
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from easyrag_b200 import _lib, encoder as enc          # noqa: E402


def timeit(fn, iters):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=147456)
    ap.add_argument("--dim", type=int, default=768)
    ap.add_argument("--iters", type=int, default=20)
    a = ap.parse_args()
    _lib.require_cuda()
    L = _lib.lib()
    dev = "cuda"
    m, d = a.rows, a.dim
    x = torch.randn(m, d, device=dev).to(torch.bfloat16)
    g = torch.ones(d, device=dev, dtype=torch.bfloat16)
    b = torch.zeros(d, device=dev, dtype=torch.bfloat16)
    out = torch.empty_like(x)
    ms = timeit(lambda: enc.layernorm(x, g, b, 1e-12, out=out), a.iters)
    print(json.dumps({"op": "layernorm", "rows": m, "dim": d, "ms": ms, "GBps": 2 * m * d * 2 / ms / 1e6}))
    ms = timeit(lambda: enc.rmsnorm(x, g, 1e-6, out=out), a.iters)
    print(json.dumps({"op": "rmsnorm", "rows": m, "dim": d, "ms": ms, "GBps": 2 * m * d * 2 / ms / 1e6}))
    hd, h, kv = 64, d // 64, max(1, d // 64 // 3)
    qkv = torch.randn(m, (h + 2 * kv) * hd, device=dev).to(torch.bfloat16)
    pos = (torch.arange(m, device=dev) % 512).to(torch.int32)
    inv = 1.0 / (1e6 ** (torch.arange(0, hd, 2, device=dev).float() / hd))
    fr = torch.arange(1024, device=dev).float()[:, None] * inv[None]
    cos, sin = fr.cos().to(torch.bfloat16).contiguous(), fr.sin().to(torch.bfloat16).contiguous()
    st = _lib.stream_ptr()
    ms = timeit(lambda: _lib.check(L.ezr_rope(_lib.ptr(qkv), qkv.stride(0), _lib.ptr(pos), _lib.ptr(cos), _lib.ptr(sin),
                                              1024, h + kv, hd, m, st), "ezr_rope"), a.iters)
    print(json.dumps({"op": "rope", "rows": m, "heads_qk": h + kv, "head_dim": hd, "ms": ms,
                      "GBps": 2 * m * (h + kv) * hd * 2 / ms / 1e6}))


if __name__ == "__main__":
    main()
