"""tcgen05.ld cost and SM-wide TMEM -> register bandwidth while the tensor core runs a tcgen05.mma chain (csrc/microbench.cu)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ai_music_generation_b200 import _C
_C.use_debug_lib()
lib = _C.lib()
out = torch.zeros(32, device="cuda", dtype=torch.int64)
iters = 400
for mma_n in (0, 64, 256):
    for x16, inflight in ((0, 1), (0, 2), (1, 1), (1, 2)):
        for nw in (4, 8, 16):
            out.zero_()
            n = mma_n if mma_n else 64
            per_mma = {64: 75, 128: 97, 256: 161}[n]
            # size the MMA chain to outlast the loads (~1.2 k cycles per load round under contention at most)
            mma_iters = 0 if mma_n == 0 else max(1, int(iters * 1500 / (4 * per_mma)))
            rc = lib.abcgpt_debug_tmem_mma_bench(out.data_ptr(), iters, nw, x16, inflight, n, mma_iters, 0)
            torch.cuda.synchronize()
            assert rc == 0
            cyc = out[2:2 + nw].float().mean().item() / iters
            nbytes = (2048 if x16 else 4096) * inflight
            mma = out[0].item() / max(1, 4 * mma_iters)
            print(f"mma N={mma_n:3d} ({mma:6.1f} cyc/MMA)  ld {'x16' if x16 else 'x32'} inflight {inflight} warps {nw:2d}: "
                  f"{cyc:7.1f} cyc per round per warp, SM-wide {nw * nbytes / cyc:6.1f} B/cyc")
