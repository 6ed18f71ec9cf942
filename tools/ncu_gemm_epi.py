"""ncu target: the epilogue-bound GEMMs of the step (GELU, GELU', the two residual adds), twice each (cfg3 shapes)."""
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
resid = torch.randn(M, C, device=dev)
xo = torch.empty(M, C, device=dev)
wp = torch.randn(C, C, device=dev).bfloat16()
for _ in range(2):
    ops.gemm(x, wfc, epilogue=ops.EPI_GELU, out=h, out2=g)                    # c_fc + GELU
    ops.gemm(dy, wproj, b_mn=True, epilogue=ops.EPI_DGELU, out=dh, aux=h)     # dgrad of mlp.c_proj + GELU'
    ops.gemm(x, wp, epilogue=ops.EPI_RESID, out=xo, aux=resid)                # attn.c_proj + residual (K = 768)
    ops.gemm(g, wproj, epilogue=ops.EPI_RESID, out=xo, aux=resid)             # mlp.c_proj + residual (K = 3072)
torch.cuda.synchronize()
print("done")
