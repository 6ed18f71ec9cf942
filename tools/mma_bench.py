"""tcgen05.mma micro-benchmark (one CTA, one issuing warp, back-to-back independent-operand MMAs into one accumulator):
cycles per 128 x N x 16 bf16 MMA by N and operand source."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ai_music_generation_b200 import _C
_C.use_debug_lib()  # instrumentation entry points live in libabcgpt_debug.so (include/abcgpt_debug.h)
lib = _C.lib()
out = torch.zeros(1, device="cuda", dtype=torch.int64)
iters = 500
names = {0: "A smem K-major", 1: "A smem MN-major", 2: "A TMEM"}
for alt in (0, 1):
  print("two alternating accumulators" if alt else "one accumulator (dependent chain)")
  for n in (64, 128, 256):
    for amode in (0, 2):
        if amode == 2 and n == 256 and alt:
            continue
        for b_mn in (0,):
            rc = lib.abcgpt_debug_mma_bench(out.data_ptr(), iters, n, amode + 4 * b_mn + 8 * alt, 0)
            torch.cuda.synchronize()
            cyc = out.item() / (4 * iters)
            ab = (4096 if amode < 2 else 0) + n * 32
            print(f"N={n:3d} {names[amode]:16s} B {'MN' if b_mn else 'K '}-major: {cyc:6.1f} cyc/MMA (math floor {n / 2:.0f}), "
                  f"smem operand bytes {ab}, {ab / cyc:5.1f} B/cyc")

print("CTA pair (cta_group::2), 256 x N x 16, SS K-major")
for n in (64, 128, 256):
    lib.abcgpt_debug_mma_bench(out.data_ptr(), iters, n, -1, 0)
    torch.cuda.synchronize()
    cyc = out.item() / (4 * iters)
    print(f"N={n:3d}: {cyc:6.1f} cyc/MMA (math floor {n / 2:.0f}); per-CTA smem operand bytes {4096 + n * 16}")
