"""Tile-shape sweep of the GEMM roles of one transformer block on a real B200: every role of the training step (forward,
dgrad, wgrad of c_attn / attn.c_proj / c_fc / mlp.c_proj) timed with each tile hint of `abcgpt_gemm_bf16` (128 / 256 = one
CTA per 128 x bn tile, 512 = CTA pair per 256 x 256 tile, 0 = the library's own choice) and, for wgrad, a few split-K
counts.  Every variant cycles through NSET private operand / output sets so the operands come from HBM as in the step.

    python tools/gemm_tile_sweep.py --tokens 16384 --embd 384            # cfg2 (baby GPT)
    python tools/gemm_tile_sweep.py --tokens 32768 --embd 768            # cfg3
Writes one JSON line per (role, variant) to stdout; `--out FILE` also appends them to FILE.
"""
from __future__ import annotations

import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    from ai_music_generation_b200 import ops

    ap = argparse.ArgumentParser()
    ap.add_argument("--tokens", type=int, default=16384)
    ap.add_argument("--embd", type=int, default=384)
    ap.add_argument("--iters", type=int, default=40)
    ap.add_argument("--hints", type=str, default="0,128,256,512")
    ap.add_argument("--out", type=str, default="")
    args = ap.parse_args()
    M, C = args.tokens, args.embd
    hints = [int(h) for h in args.hints.split(",")]
    dev = "cuda"
    torch.manual_seed(0)

    def bf(*shape):
        return (torch.randn(*shape, device=dev) * 0.5).bfloat16()

    # (name, builder) — builder(nset) returns a list of callables taking (hint, splits)
    roles = []

    def add_fwd(name, N, K, epi):
        def build(nset):
            sets = []
            for _ in range(nset):
                a, w = bf(M, K), bf(N, K)
                if epi == ops.EPI_GELU:
                    o, o2 = torch.empty(M, N, device=dev, dtype=torch.bfloat16), torch.empty(M, N, device=dev, dtype=torch.bfloat16)
                    sets.append(lambda h, s, a=a, w=w, o=o, o2=o2: ops.gemm(a, w, epilogue=epi, out=o, out2=o2, tile_n=h))
                elif epi == ops.EPI_RESID:
                    x, o = torch.randn(M, N, device=dev), torch.empty(M, N, device=dev)
                    sets.append(lambda h, s, a=a, w=w, o=o, x=x: ops.gemm(a, w, epilogue=epi, out=o, aux=x, tile_n=h))
                else:
                    o = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
                    sets.append(lambda h, s, a=a, w=w, o=o: ops.gemm(a, w, epilogue=epi, out=o, tile_n=h))
            return sets
        roles.append((name, 2.0 * M * N * K, build, False))

    def add_dgrad(name, N_out, K_in, epi):
        # dX[M, K_in] = dY[M, N_out] W[N_out, K_in]  (B = W stored [N_out, K_in] = MN-major for this product)
        def build(nset):
            sets = []
            for _ in range(nset):
                dy, w = bf(M, N_out), bf(N_out, K_in)
                o = torch.empty(M, K_in, device=dev, dtype=torch.bfloat16)
                if epi == ops.EPI_DGELU:
                    h = bf(M, K_in)
                    sets.append(lambda hn, s, dy=dy, w=w, o=o, h=h: ops.gemm(dy, w, b_mn=True, epilogue=epi, out=o, aux=h, tile_n=hn))
                else:
                    sets.append(lambda hn, s, dy=dy, w=w, o=o: ops.gemm(dy, w, b_mn=True, epilogue=epi, out=o, tile_n=hn))
            return sets
        roles.append((name, 2.0 * M * N_out * K_in, build, False))

    def add_wgrad(name, N_out, K_in):
        # dW[N_out, K_in] += dY[M, N_out]^T X[M, K_in]
        def build(nset):
            sets = []
            for _ in range(nset):
                dy, x = bf(M, N_out), bf(M, K_in)
                o = torch.zeros(N_out, K_in, device=dev)
                sets.append(lambda hn, s, dy=dy, x=x, o=o: ops.gemm(dy, x, a_mn=True, b_mn=True, epilogue=ops.EPI_F32_RED, out=o,
                                                                   tile_n=hn, splits=s))
            return sets
        roles.append((name, 2.0 * M * N_out * K_in, build, True))

    add_fwd("fwd c_attn", 3 * C, C, ops.EPI_BF16)
    add_fwd("fwd attn.c_proj", C, C, ops.EPI_BF16)
    add_fwd("fwd c_fc+GELU", 4 * C, C, ops.EPI_GELU)
    add_fwd("fwd mlp.c_proj+resid", C, 4 * C, ops.EPI_RESID)
    add_dgrad("dgrad mlp.c_proj+GELU'", C, 4 * C, ops.EPI_DGELU)
    add_dgrad("dgrad c_fc", 4 * C, C, ops.EPI_BF16)
    add_dgrad("dgrad attn.c_proj", C, C, ops.EPI_BF16)
    add_dgrad("dgrad c_attn", 3 * C, C, ops.EPI_BF16)
    add_wgrad("wgrad mlp.c_proj", C, 4 * C)
    add_wgrad("wgrad c_fc", 4 * C, C)
    add_wgrad("wgrad attn.c_proj", C, C)
    add_wgrad("wgrad c_attn", 3 * C, C)

    fout = open(args.out, "a") if args.out else None
    for name, flops, build, is_wgrad in roles:
        # enough private sets that one pass over them exceeds the 126 MB L2
        bytes_per_set = max(1, int(flops / (2.0 * min(M, C)) * 2))  # rough: the largest operand
        nset = max(2, min(16, (300 << 20) // max(bytes_per_set, 1 << 20)))
        sets = build(nset)
        variants = [(h, 0) for h in hints]
        if is_wgrad:
            variants += [(h, s) for h in hints if h != 0 for s in (2, 4, 8, 16)]
        for h, s in variants:
            try:
                for f in sets:
                    f(h, s)
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                n = 0
                for _ in range(max(1, args.iters // nset)):
                    for f in sets:
                        f(h, s)
                        n += 1
                e1.record()
                torch.cuda.synchronize()
                ms = e0.elapsed_time(e1) / n
                rec = {"role": name, "tile": h, "splits": s, "ms": round(ms, 5), "tflops": round(flops / ms / 1e9, 1)}
            except Exception as ex:  # an unsupported combination is a result too
                rec = {"role": name, "tile": h, "splits": s, "error": str(ex)[:120]}
            line = json.dumps(rec)
            print(line, flush=True)
            if fout:
                fout.write(line + "\n")
        del sets
        torch.cuda.empty_cache()
    if fout:
        fout.close()


if __name__ == "__main__":
    main()
