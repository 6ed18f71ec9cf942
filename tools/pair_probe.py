"""Operand-format probe of the CTA-pair MMA (tcgen05.mma.cta_group::2) as the pair attention kernels use it:
D[256 x 64] = A[256 x 64] B[64 x 64], B split by N between the two CTAs — K-major row halves (SWIZZLE_128B) and MN-major column
halves (64-byte rows, SWIZZLE_64B) — with A from shared memory and from tensor memory.  Prints the max error per mode."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ai_music_generation_b200 import _C
_C.use_debug_lib()
lib = _C.lib()
torch.manual_seed(0)
a = torch.randn(256, 64, device="cuda").bfloat16()
b = torch.randn(64, 64, device="cuda").bfloat16()      # B[k][n]
ref = a.float() @ b.float()
for mode, name in ((1, "B K-major row halves, A smem"), (0, "B MN-major SW64 column halves, A smem"),
                   (3, "B K-major row halves, A TMEM"), (2, "B MN-major SW64 column halves, A TMEM")):
    d = torch.full((256, 64), float("nan"), device="cuda")
    barg = b.t().contiguous() if mode & 1 else b
    rc = lib.abcgpt_debug_pair_probe(a.data_ptr(), barg.data_ptr(), d.data_ptr(), mode, 0)
    torch.cuda.synchronize()
    err = (d - ref).abs()
    bad = (err > 0.05) | ~torch.isfinite(d)
    print(f"mode {mode} ({name}): rc {rc} max err {err[torch.isfinite(err)].max().item() if torch.isfinite(err).any() else float('nan'):.4f} "
          f"bad {int(bad.sum())} / {d.numel()}  rows-bad {int(bad.any(1).sum())} cols-bad {bad.any(0).int().tolist() if bad.any() else []}")
