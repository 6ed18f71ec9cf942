"""ncu target: the two epilogue-bound GEMMs of the step, once each (cfg3 shapes)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ai_music_generation_b200 import ops
M, C = 32768, 768
dev = "cuda"
torch.manual_seed(0)
x = torch.randn(M, C, device=dev).bfloat16()
wfc = torch.randn(4 * C, C, device=dev).bfloat16()
h = torch.empty(M, 4 * C, device=dev, dtype=torch.bfloat16)
g = torch.empty(M, 4 * C, device=dev, dtype=torch.bfloat16)
dy = torch.randn(M, C, device=dev).bfloat16()
wproj = torch.randn(C, 4 * C, device=dev).bfloat16()
dh = torch.empty(M, 4 * C, device=dev, dtype=torch.bfloat16)
for _ in range(2):
    ops.gemm(x, wfc, epilogue=ops.EPI_GELU, out=h, out2=g)
    ops.gemm(dy, wproj, b_mn=True, epilogue=ops.EPI_DGELU, out=dh, aux=h)
torch.cuda.synchronize()
print("done")
