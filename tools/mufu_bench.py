"""MUFU.EX2 issue rate on one SM by warps per scheduler: cycles per warp instruction (tools for DESIGN.md 4.0)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ai_music_generation_b200 import _C
_C.use_debug_lib()
lib = _C.lib()
out = torch.zeros(64, device="cuda", dtype=torch.int64)
sink = torch.zeros(1, device="cuda")
iters = 2000
for mode, name in ((0, "ex2 only"), (3, "ex2.f16x2 (2 per instr)"), (4, "ex2.bf16x2 (2 per instr)"), (1, "fma + ex2"), (2, "FMA-pipe exp2 (cubic)")):
    for warps in (4, 8, 16, 32):
        lib.abcgpt_debug_mufu_bench(out.data_ptr(), sink.data_ptr(), iters, warps, mode, 0)
        torch.cuda.synchronize()
        cyc = out[:warps].float().mean().item()
        per = cyc / (iters * 16)
        print(f"{name:24s} {warps:2d} warps ({warps // 4}/scheduler): {per:6.2f} cycles per element-instruction per warp, "
              f"{warps / 4 / per * 32:6.1f} elements/clk/scheduler")
