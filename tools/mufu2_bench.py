"""MUFU.EX2 throughput by operand format on one SM (csrc/microbench.cu mufu2_bench_kernel, compile-time modes)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ai_music_generation_b200 import _C
_C.use_debug_lib()
lib = _C.lib()
out = torch.zeros(64, device="cuda", dtype=torch.int64)
sink = torch.zeros(1, device="cuda")
iters = 2000
MODES = ((0, "ex2.f32", 16, 16), (1, "ex2.f16x2", 16, 32), (2, "ex2.bf16x2", 16, 32), (3, "ffma2 + 2 ex2.f32 + cvt.bf16x2 (per pair)", 16, 16),
         (4, "ffma2 + cvt.f16x2 + ex2.f16x2 (per pair)", 8, 16))
for mode, name, mufu_per_iter, elems_per_iter in MODES:
    for warps in (4, 8, 16, 32):
        lib.abcgpt_debug_mufu2_bench(out.data_ptr(), sink.data_ptr(), iters, warps, mode, 0)
        torch.cuda.synchronize()
        cyc = out[:warps].float().max().item()
        print(f"{name:44s} {warps:2d} warps ({warps // 4}/scheduler): {cyc / (iters * mufu_per_iter):6.2f} cycles per MUFU instruction per warp, "
              f"{warps * 32 * elems_per_iter * iters / cyc:6.1f} exponentials/clk/SM")
