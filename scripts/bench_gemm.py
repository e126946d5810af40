#!/usr/bin/env python
"""Per-shape throughput of the encoder GEMM kernel (ezr_gemm_bf16): the four projections of a BERT-base layer and the
Qwen2-shaped ones at the token counts the encode bench runs (512 sequences of U[64,512] tokens ~ 147k rows).

    python scripts/bench_gemm.py [--rows 147456]      # one JSON line per shape
"""
import argparse
import json
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from easyrag_b200 import _lib, encoder as enc          # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=147456)
    ap.add_argument("--iters", type=int, default=20)
    a = ap.parse_args()
    _lib.require_cuda()
    dev = "cuda"
    g = torch.Generator(device=dev).manual_seed(0)
    shapes = [("qkv", 768, 2304, enc.EPI_NONE, True, False), ("attn_out", 768, 768, enc.EPI_NONE, True, True),
              ("ffn_up_gelu", 768, 3072, enc.EPI_GELU, True, False), ("ffn_down", 3072, 768, enc.EPI_NONE, True, True),
              ("ffn_up_nogelu", 768, 3072, enc.EPI_NONE, True, False), ("gate_up_swiglu", 768, 6144, enc.EPI_SWIGLU, False, False),
              ("square_4k", 4096, 4096, enc.EPI_NONE, False, False)]
    for name, k, n, epi, bias, res in shapes:
        m = a.rows
        x = (torch.randn(m, k, generator=g, device=dev) * 0.5).to(torch.bfloat16)
        w = (torch.randn(n, k, generator=g, device=dev) * 0.05).to(torch.bfloat16)
        b = (torch.randn(n, generator=g, device=dev) * 0.1).to(torch.bfloat16) if bias else None
        n_out = n // 2 if epi == enc.EPI_SWIGLU else n
        r = (torch.randn(m, n_out, generator=g, device=dev) * 0.5).to(torch.bfloat16) if res else None
        out = torch.empty(m, n_out, dtype=torch.bfloat16, device=dev)
        for _ in range(3):
            enc.gemm(x, w, bias=b, residual=r, out=out, epilogue=epi)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(a.iters):
            enc.gemm(x, w, bias=b, residual=r, out=out, epilogue=epi)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / a.iters
        print(json.dumps({"gemm": name, "M": m, "N": n, "K": k, "epilogue": epi, "ms": ms,
                          "tflops": 2.0 * m * n * k / (ms * 1e-3) / 1e12}))


if __name__ == "__main__":
    main()
