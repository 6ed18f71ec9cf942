"""Where do the microseconds of a SMALL pair-kernel GEMM go?  %globaltimer stamps over all CTAs of three back-to-back launches
(debug counters of libabcgpt_debug.so): kernel entry, prologue done, first operands landed, last MMA committed, epilogues done,
exit — and the gap between one launch's exit and the next one's entry.

    python tools/gemm_timeline.py [--tokens 16384] [--embd 384] [--pdl 0|1]
"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ai_music_generation_b200 import _C, ops  # noqa: E402
_C.use_debug_lib()

ap = argparse.ArgumentParser()
ap.add_argument("--tokens", type=int, default=16384)
ap.add_argument("--embd", type=int, default=384)
ap.add_argument("--pdl", type=int, default=0)
ap.add_argument("--burn", type=int, default=0, help="back-to-back launches before the stamped ones (sustained-load clock)")
args = ap.parse_args()
M, C = args.tokens, args.embd
dev = "cuda"
BIG = (1 << 63) - 1
shapes = [("fwd c_attn", M, 3 * C, C, False, False, ops.EPI_BF16), ("fwd attn.c_proj", M, C, C, False, False, ops.EPI_BF16),
          ("fwd c_fc+GELU", M, 4 * C, C, False, False, ops.EPI_GELU), ("fwd mlp.c_proj+resid", M, C, 4 * C, False, False, ops.EPI_RESID),
          ("dgrad c_fc", M, C, 4 * C, False, True, ops.EPI_BF16), ("wgrad c_fc", 4 * C, C, M, True, True, ops.EPI_F32_RED),
          ("wgrad attn.c_proj", C, C, M, True, True, ops.EPI_F32_RED)]
ops.set_pdl(bool(args.pdl))
for name, m, n, k, amn, bmn, epi in shapes:
    A = torch.randn((k, m) if amn else (m, k), device=dev).bfloat16()
    B = torch.randn((k, n) if bmn else (n, k), device=dev).bfloat16()
    odt = torch.float32 if epi in (ops.EPI_RESID, ops.EPI_F32_RED) else torch.bfloat16
    out = torch.zeros(m, n, device=dev, dtype=odt)
    out2 = torch.zeros(m, n, device=dev, dtype=torch.bfloat16) if epi == ops.EPI_GELU else None
    aux = None
    if epi == ops.EPI_RESID:
        aux = torch.randn(m, n, device=dev)
    run = lambda: ops.gemm(A, B, a_mn=amn, b_mn=bmn, epilogue=epi, out=out, out2=out2, aux=aux, tile_n=512)
    for _ in range(3):
        run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        run()
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / 20 * 1e3
    stats = torch.zeros(3, 16, device=dev, dtype=torch.int64)
    stats[:, 8] = BIG
    stats[:, 10] = BIG
    stats[:, 15] = BIG
    torch.cuda.synchronize()
    for _ in range(args.burn):
        run()
    for i in range(3):
        _C.lib().abcgpt_debug_gemm_stats(stats[i].data_ptr())
        run()
    _C.lib().abcgpt_debug_gemm_stats(0)
    torch.cuda.synchronize()
    s = stats.tolist()
    r = s[1]
    t0 = r[8]
    rel = lambda v: (v - t0) / 1e3
    print(f"{name:22s} {m}x{n}x{k}: {us:6.1f} us per back-to-back launch | entry 0, prologue done {rel(r[15]):5.1f}..{rel(r[9]):5.1f}, first operands "
          f"{rel(r[10]):5.1f}..{rel(r[11]):5.1f}, last MMA {rel(r[12]):5.1f}, epilogues {rel(r[13]):5.1f}, exit {rel(r[14]):5.1f} us; "
          f"SM clock {r[6] / max(r[7], 1):4.2f} GHz, {2.0 * m * n * k / max(r[14] - r[8], 1) / 1e3:6.0f} TF/s inside; gap to next entry {(s[2][8] - r[14]) / 1e3:5.1f} us, previous exit -> this entry {(r[8] - s[0][14]) / 1e3:5.1f} us")
