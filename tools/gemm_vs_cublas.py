"""The pair GEMM against the library GEMM (torch.matmul -> cuBLASLt, bf16 in / bf16 out) on the GEMM shapes of one transformer
block, under the SAME sustained load: each variant runs `--burst` back-to-back launches per timing, variants alternate
(A/B/A/B), so both see the clock the power cap allows.  cuBLAS computes the plain product only (no GELU / residual / fp32
accumulation into a gradient): the comparison is conservative for the fused roles.

    python tools/gemm_vs_cublas.py [--tokens 32768] [--embd 768] [--burst 200] [--rounds 3]
"""
from __future__ import annotations

import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ai_music_generation_b200 import ops  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--tokens", type=int, default=32768)
ap.add_argument("--embd", type=int, default=768)
ap.add_argument("--burst", type=int, default=200)
ap.add_argument("--rounds", type=int, default=3)
args = ap.parse_args()
M, C = args.tokens, args.embd
dev = "cuda"
torch.manual_seed(0)


def bf(*s):
    return (torch.randn(*s, device=dev) * 0.5).bfloat16()


def timed(fn, n):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


roles = []
# forward: Y[M, N] = X[M, K] W[N, K]^T
for name, N, K, epi in (("fwd c_attn", 3 * C, C, ops.EPI_BF16), ("fwd attn.c_proj", C, C, ops.EPI_BF16),
                        ("fwd c_fc (+GELU ours)", 4 * C, C, ops.EPI_GELU), ("fwd mlp.c_proj (+resid ours)", C, 4 * C, ops.EPI_RESID)):
    x, w = bf(M, K), bf(N, K)
    yl = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    if epi == ops.EPI_GELU:
        o, o2 = torch.empty_like(yl), torch.empty_like(yl)
        ours = lambda x=x, w=w, o=o, o2=o2: ops.gemm(x, w, epilogue=ops.EPI_GELU, out=o, out2=o2)
    elif epi == ops.EPI_RESID:
        r, o = torch.randn(M, N, device=dev), torch.empty(M, N, device=dev)
        ours = lambda x=x, w=w, o=o, r=r: ops.gemm(x, w, epilogue=ops.EPI_RESID, out=o, aux=r)
    else:
        o = torch.empty_like(yl)
        ours = lambda x=x, w=w, o=o: ops.gemm(x, w, epilogue=ops.EPI_BF16, out=o)
    lib = lambda x=x, w=w, yl=yl: torch.matmul(x, w.t(), out=yl)
    roles.append((name, 2.0 * M * N * K, ours, lib))
# dgrad: dX[M, K] = dY[M, N] W[N, K]
for name, N, K, epi in (("dgrad c_fc", 4 * C, C, ops.EPI_BF16), ("dgrad mlp.c_proj (+GELU' ours)", C, 4 * C, ops.EPI_DGELU),
                        ("dgrad c_attn", 3 * C, C, ops.EPI_BF16)):
    dy, w = bf(M, N), bf(N, K)
    o, yl = torch.empty(M, K, device=dev, dtype=torch.bfloat16), torch.empty(M, K, device=dev, dtype=torch.bfloat16)
    if epi == ops.EPI_DGELU:
        h = bf(M, K)
        ours = lambda dy=dy, w=w, o=o, h=h: ops.gemm(dy, w, b_mn=True, epilogue=ops.EPI_DGELU, out=o, aux=h)
    else:
        ours = lambda dy=dy, w=w, o=o: ops.gemm(dy, w, b_mn=True, epilogue=ops.EPI_BF16, out=o)
    lib = lambda dy=dy, w=w, yl=yl: torch.matmul(dy, w, out=yl)
    roles.append((name, 2.0 * M * N * K, ours, lib))
# wgrad: dW[N, K] += dY[M, N]^T X[M, K]
for name, N, K in (("wgrad c_fc (fp32 += ours)", 4 * C, C), ("wgrad c_attn (fp32 += ours)", 3 * C, C)):
    dy, x = bf(M, N), bf(M, K)
    o, yl = torch.zeros(N, K, device=dev), torch.empty(N, K, device=dev, dtype=torch.bfloat16)
    ours = lambda dy=dy, x=x, o=o: ops.gemm(dy, x, a_mn=True, b_mn=True, epilogue=ops.EPI_F32_RED, out=o)
    lib = lambda dy=dy, x=x, yl=yl: torch.matmul(dy.t(), x, out=yl)
    roles.append((name, 2.0 * M * N * K, ours, lib))

for name, flops, ours, lib in roles:
    for f in (ours, lib):
        timed(f, 20)
    t_ours, t_lib = [], []
    for _ in range(args.rounds):
        t_ours.append(timed(ours, args.burst))
        t_lib.append(timed(lib, args.burst))
    a, b = min(t_ours), min(t_lib)
    print(json.dumps({"role": name, "ours_ms": round(a, 4), "ours_tflops": round(flops / a / 1e9, 1), "cublas_ms": round(b, 4),
                      "cublas_tflops": round(flops / b / 1e9, 1), "ours_over_cublas": round(b / a, 3)}), flush=True)
