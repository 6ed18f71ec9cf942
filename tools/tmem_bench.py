"""tcgen05.ld micro-benchmark: cycles per 32x32b.x32 load (4 KB per warp) for 1..8 warps and 1/2/4 loads in flight."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ai_music_generation_b200 import _C
_C.use_debug_lib()  # instrumentation entry points live in libabcgpt_debug.so (include/abcgpt_debug.h)
lib = _C.lib()
out = torch.zeros(8, device="cuda", dtype=torch.int64)
iters = 2000
for inflight in (1, 2, 4):
    for nw in (1, 2, 4, 8):
        out.zero_()
        rc = lib.abcgpt_debug_tmem_ld_bench(out.data_ptr(), iters, nw, inflight, 0)
        torch.cuda.synchronize()
        cyc = out[:nw].float().mean().item() / iters
        print(f"inflight {inflight} warps {nw}: {cyc:7.1f} cyc/iter, {cyc / inflight:6.1f} cyc per 4 KB load per warp, "
              f"SM-wide {nw * inflight * 4096 / cyc:6.1f} B/cyc")
