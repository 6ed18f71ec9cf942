"""Does a small kernel on a second stream slow the persistent pair GEMM down?  One GPU: a loop of cfg3 dgrad GEMMs on the main
stream, timed alone and with the gradient-exchange kernel (world = 1: plain loads / stores, same code and footprint as the NVLS
form) looping over a 497 MB arena on a high-priority side stream, for several launch shapes of the side kernel.

    python tools/coresidency_probe.py
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ai_music_generation_b200 import ops  # noqa: E402

dev = "cuda"
M, C = 32768, 768
dy = (torch.randn(M, 4 * C, device=dev) * 0.5).bfloat16()
w = (torch.randn(4 * C, C, device=dev) * 0.5).bfloat16()
out = torch.empty(M, C, device=dev, dtype=torch.bfloat16)
arena = torch.randn(124_000_000, device=dev)
table = torch.zeros(65536, device=dev)
side = torch.cuda.Stream(priority=-1)
NG = 60


def gemm_loop():
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(NG):
        ops.gemm(dy, w, b_mn=True, epilogue=ops.EPI_BF16, out=out)
    e1.record()
    return e0, e1


for _ in range(2):
    gemm_loop()
torch.cuda.synchronize()
e0, e1 = gemm_loop()
torch.cuda.synchronize()
base = e0.elapsed_time(e1) / NG
print(f"GEMM alone: {base * 1e3:.1f} us per launch")
for blocks, threads in ((148, 128), (16, 256), (32, 128), (8, 512), (32, 256)):
    # side kernel alone
    torch.cuda.synchronize()
    with torch.cuda.stream(side):
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s0.record()
        for _ in range(4):
            ops.nvls_allreduce_sumsq(arena.data_ptr(), arena.numel(), 0, 1, table.data_ptr(), blocks, threads)
        s1.record()
    torch.cuda.synchronize()
    side_alone = s0.elapsed_time(s1) / 4
    # both
    ready = torch.cuda.Event()
    ready.record()
    with torch.cuda.stream(side):
        side.wait_event(ready)
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s0.record()
        nside = max(1, int(base * NG / side_alone * 0.9))
        for _ in range(nside):
            ops.nvls_allreduce_sumsq(arena.data_ptr(), arena.numel(), 0, 1, table.data_ptr(), blocks, threads)
        s1.record()
    e0, e1 = gemm_loop()
    torch.cuda.synchronize()
    both = e0.elapsed_time(e1) / NG
    print(f"side {blocks:4d} x {threads:3d}: side alone {side_alone:6.3f} ms per pass ({arena.numel() * 8 / side_alone / 1e9:5.2f} TB/s), "
          f"{nside} passes beside the GEMMs: {s0.elapsed_time(s1) / nside:6.3f} ms per pass; GEMM {both * 1e3:6.1f} us per launch "
          f"(+{(both / base - 1) * 100:4.1f} %)")
