#!/usr/bin/env python
"""Encoder-side measurement for BASELINE.json configs[1]: "GTE-base 768-d encode + cosine top-10, 100k chunks,
1k queries, 1xB200" (tensor-bound; report TFLOP/s vs the measured bf16 peak).

    python bench_encode.py [--arch bert|qwen2] [--chunks N] [--batch 128]

Random-init weights of the named architecture (no checkpoints offline), synthetic token ids, chunk length
U[64,512], query length U[8,48].  Prints one JSON line; not the driver's bench (that is bench.py).
"""
from __future__ import annotations

import argparse
import json
import sys
import time
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))


def main(from_bench=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--arch", default="bert", choices=["bert", "qwen2"])
    ap.add_argument("--chunks", type=int, default=20_000)
    ap.add_argument("--queries", type=int, default=1_000)
    ap.add_argument("--batch", type=int, default=128)          # embed_batch_size, pipeline.py:105
    ap.add_argument("--layers", type=int, default=12)
    ap.add_argument("--dim", type=int, default=768)
    ap.add_argument("--steps", type=int, default=3)
    args = ap.parse_known_args()[0]
    from easyrag_b200 import _lib, batched
    from easyrag_b200.encoder import (BertConfig, BertEncoder, PackedBatch, Qwen2Config, Qwen2Encoder, random_state)
    from easyrag_b200.index import DenseIndex
    _lib.require_cuda()
    L = _lib.lib()
    dev = torch.device("cuda", 0)
    d = args.dim
    if args.arch == "bert":
        cfg = BertConfig(vocab_size=21128, hidden_size=d, intermediate_size=4 * d, num_hidden_layers=args.layers,
                         num_attention_heads=d // 64, max_position_embeddings=512)
        model = BertEncoder(cfg, random_state("bert", cfg, 1), device=dev)
    else:
        cfg = Qwen2Config(vocab_size=151646, hidden_size=d, intermediate_size=4 * d, num_hidden_layers=args.layers,
                          num_attention_heads=d // 64, num_key_value_heads=max(1, d // 64 // 3),
                          max_position_embeddings=1024)
        model = Qwen2Encoder(cfg, random_state("qwen2", cfg, 1), device=dev)
    g = torch.Generator().manual_seed(2)
    lens = torch.randint(64, 513, (args.chunks,), generator=g).tolist()
    qlens = torch.randint(8, 49, (args.queries,), generator=g).tolist()

    def make_batches(ls):
        out = []
        for i in range(0, len(ls), args.batch):
            part = ls[i:i + args.batch]
            ids = torch.randint(1, cfg.vocab_size, (sum(part),), generator=g, dtype=torch.int32)
            cu = torch.tensor([0] + list(torch.tensor(part).cumsum(0)), dtype=torch.int32)
            pos = torch.cat([torch.arange(n, dtype=torch.int32) for n in part])
            out.append((PackedBatch(ids=ids.to(dev), cu=cu.to(dev), positions=pos.to(dev), max_len=max(part),
                                    n_seq=len(part)), part))
        return out

    cb, qb = make_batches(lens), make_batches(qlens)
    corpus = torch.empty(args.chunks, d, dtype=torch.bfloat16, device=dev)

    def encode_all(batches, out=None):
        o = 0
        for b, part in batches:
            eb, _ = model.embed_packed(b)
            if out is not None:
                out[o:o + len(part)] = eb
            o += len(part)

    encode_all(cb[:4])
    torch.cuda.synchronize()
    _lib.check(L.ezr_profile_reset())
    _lib.check(L.ezr_profile_enable(1))
    t0 = time.perf_counter()
    encode_all(cb, corpus)
    torch.cuda.synchronize()
    t_corpus = time.perf_counter() - t0
    _lib.check(L.ezr_profile_enable(0))
    prof = {n: _lib.profile_read(n) for n in ("enc_gemm", "enc_attn", "enc_other")}
    flops = model.flops(lens)
    # queries: encode + cosine top-10 against the corpus just produced
    index = DenseIndex(corpus, device=dev)
    qv = torch.empty(args.queries, d, dtype=torch.bfloat16, device=dev)
    encode_all(qb, qv)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        encode_all(qb, qv)
        res = batched.dense_topk(index, qv, 10)
    torch.cuda.synchronize()
    t_query = (time.perf_counter() - t0) / args.steps
    peaks = {}
    pk = ROOT / "MEASURED_PEAKS.json"
    if pk.exists():
        peaks = json.loads(pk.read_text())
    peak = float(peaks.get("bf16_tflops_sustained", 1400.0))
    gemm_ms = prof["enc_gemm"][0]
    # GEMM-only flops: everything except the 4 L^2 d attention term
    attn_flops = sum(cfg.num_hidden_layers * 4 * n * n * d for n in lens)
    line = {
        "bench": "encode", "arch": args.arch, "layers": args.layers, "dim": d, "chunks": args.chunks,
        "tokens": sum(lens), "batch": args.batch, "encode_s": t_corpus, "chunks_per_s": args.chunks / t_corpus,
        "tokens_per_s": sum(lens) / t_corpus, "model_tflops": flops / t_corpus / 1e12,
        "gemm": {"ms": gemm_ms, "launches": prof["enc_gemm"][1],
                 "tflops": (flops - attn_flops) / (gemm_ms * 1e-3) / 1e12 if gemm_ms else None,
                 "frac_of_measured_bf16_sustained": (flops - attn_flops) / (gemm_ms * 1e-3) / 1e12 / peak if gemm_ms else None},
        "attention": {"ms": prof["enc_attn"][0], "tflops": attn_flops / (prof["enc_attn"][0] * 1e-3) / 1e12 if prof["enc_attn"][0] else None},
        "other_ms": prof["enc_other"][0],
        "queries": {"n": args.queries, "encode_plus_top10_s": t_query, "queries_per_s": args.queries / t_query},
        "peak_tflops": peak, "dtype": "bf16", "data": "synthetic, random-init weights",
    }
    print(json.dumps(line))


if __name__ == "__main__":
    main()
