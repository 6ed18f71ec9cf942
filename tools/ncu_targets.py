"""Small fixed launch sequence for `ncu --set full` captures (one GPU, a handful of launches).

    python tools/ncu_targets.py [gemm] [attn] [ln] [adamw]
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ai_music_generation_b200 import ops  # noqa: E402

which = set(a for a in sys.argv[1:] if a != "once") or {"gemm", "attn", "ln", "adamw"}
REP = 1 if "once" in sys.argv[1:] else 2  # "once": every kernel exactly one time (small ncu reports)
dev = "cuda"
torch.manual_seed(0)
M, C, H, B, T = 32768, 768, 12, 32, 1024
if "gemm" in which:
    x = torch.randn(M, C, device=dev).bfloat16()
    w = torch.randn(3 * C, C, device=dev).bfloat16()
    out = torch.empty(M, 3 * C, device=dev, dtype=torch.bfloat16)
    for _ in range(REP):
        ops.gemm(x, w, epilogue=ops.EPI_BF16, out=out)                       # fwd NT, N=2304
    dy = torch.randn(M, 4 * C, device=dev).bfloat16()
    wfc = torch.randn(4 * C, C, device=dev).bfloat16()
    dx = torch.empty(M, C, device=dev, dtype=torch.bfloat16)
    ops.gemm(dy, wfc, b_mn=True, epilogue=ops.EPI_BF16, out=dx)             # dgrad, K=3072
    dw = torch.zeros(4 * C, C, device=dev)
    ops.gemm(dy, x, a_mn=True, b_mn=True, epilogue=ops.EPI_F32_RED, out=dw)  # wgrad
    resid = torch.randn(M, C, device=dev)
    xo = torch.empty(M, C, device=dev)
    wp = torch.randn(C, C, device=dev).bfloat16()
    ops.gemm(x, wp, epilogue=ops.EPI_RESID, out=xo, aux=resid)              # c_proj + residual
    h = torch.empty(M, 4 * C, device=dev, dtype=torch.bfloat16)
    g = torch.empty(M, 4 * C, device=dev, dtype=torch.bfloat16)
    ops.gemm(x, wfc, epilogue=ops.EPI_GELU, out=h, out2=g)                   # c_fc + GELU (two TMA-store streams)
    wpr = torch.randn(C, 4 * C, device=dev).bfloat16()
    dh = torch.empty(M, 4 * C, device=dev, dtype=torch.bfloat16)
    ops.gemm(x, wpr, b_mn=True, epilogue=ops.EPI_DGELU, out=dh, aux=h)       # dgrad of mlp.c_proj + GELU'
    ops.gemm(g, wpr, epilogue=ops.EPI_RESID, out=xo, aux=resid)              # mlp.c_proj + residual (K = 3072)
if "attn" in which:
    qkv = torch.randn(B * T, 3 * C, device=dev).bfloat16()
    o = torch.empty(B * T, C, device=dev, dtype=torch.bfloat16)
    lse = torch.empty(B, H, T, device=dev)
    do = torch.randn(B * T, C, device=dev).bfloat16()
    dqkv = torch.empty(B * T, 3 * C, device=dev, dtype=torch.bfloat16)
    delta = torch.empty(B, H, T, device=dev)
    for _ in range(REP):
        ops.attn_fwd(qkv, o, lse, B, T, H)
        ops.attn_bwd(qkv, o, do, lse, delta, dqkv, B, T, H)
if "ln" in which:
    x = torch.randn(M, C, device=dev)
    w = torch.ones(C, device=dev)
    y = torch.empty(M, C, device=dev, dtype=torch.bfloat16)
    st = torch.empty(2, M, device=dev)
    dy = torch.randn(M, C, device=dev).bfloat16()
    dres = torch.randn(M, C, device=dev)
    dxo = torch.empty(M, C, device=dev)
    dxb = torch.empty(M, C, device=dev, dtype=torch.bfloat16)
    dw = torch.zeros(C, device=dev)
    for _ in range(REP):
        ops.layernorm_fwd(x, w, None, y, st[0], st[1])
        ops.layernorm_fwd_resid(x, dy, dxo, w, None, y, st[0], st[1])   # residual add fused in front of the normalisation
        ops.layernorm_bwd(dy, x, w, st[0], st[1], dres, dxo, dxb, dw, None)
if "adamw" in which:
    n = 85813248
    p = torch.randn(n, device=dev)
    g = torch.randn(n, device=dev)
    m = torch.zeros(n, device=dev)
    v = torch.zeros(n, device=dev)
    sh = torch.empty(n, device=dev, dtype=torch.bfloat16)
    ss = torch.zeros(1, device=dev)
    for _ in range(REP):
        ops.sumsq(g, ss)
        ops.adamw(p, g, m, v, sh, lr=1e-3, beta1=0.9, beta2=0.95, eps=1e-8, weight_decay=0.1, step=3, sumsq=ss, max_norm=1.0)
torch.cuda.synchronize()
print("ncu_targets done")
