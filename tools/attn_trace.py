"""Phase timeline of one attention-forward CTA (the 8-tile row block of head 0): cycles between stamps."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ai_music_generation_b200 import _C, ops
_C.use_debug_lib()  # instrumentation entry points live in libabcgpt_debug.so (include/abcgpt_debug.h)
B, T, H = 32, 1024, 12
C = H * 64
qkv = torch.randn(B * T, 3 * C, device="cuda").bfloat16()
o = torch.empty(B * T, C, device="cuda", dtype=torch.bfloat16)
lse = torch.empty(B, H, T, device="cuda")
for _ in range(3):
    ops.attn_fwd(qkv, o, lse, B, T, H)
tr = torch.zeros(128 * 8, device="cuda", dtype=torch.int64)
_C.lib().abcgpt_debug_attn_trace(tr.data_ptr())
ops.attn_fwd(qkv, o, lse, B, T, H)
torch.cuda.synchronize()
_C.lib().abcgpt_debug_attn_trace(0)

t = tr.view(-1, 8).cpu()
n = int((t[:, 4] > 0).sum().item())
print("step item tile: wait_S  compute  wait_Pbuf  store+arrive | total | since previous step start")
for j in range(n):
    r = t[j]
    print(j, int(r[5]), int(r[6]), ":", (r[1] - r[0]).item(), (r[2] - r[1]).item(), (r[3] - r[2]).item(), (r[4] - r[3]).item(), "|", (r[4] - r[0]).item(),
          "|", (r[0] - t[j - 1][0]).item() if j else 0)
cyc, ns = (t[n - 1][0] - t[0][0]).item(), (t[n - 1][7] - t[0][7]).item()
print(f"forward CTA 0: {n} steps, {cyc} cycles, {ns} ns -> {cyc / max(ns, 1) * 1e3:.0f} MHz effective SM clock, {cyc / (n - 1):.0f} cycles per step")

# backward kernels: stamps of group 0's first 64 own steps (across items): dq at [0,512), dkv at [512,1024)
do = torch.randn(B * T, C, device="cuda").bfloat16()
dqkv = torch.empty(B * T, 3 * C, device="cuda", dtype=torch.bfloat16)
delta = torch.empty(B, H, T, device="cuda")
for _ in range(2):
    ops.attn_bwd(qkv, o, do, lse, delta, dqkv, B, T, H)
tr2 = torch.zeros(2048, device="cuda", dtype=torch.int64)
_C.lib().abcgpt_debug_attn_trace(tr2.data_ptr())
ops.attn_bwd(qkv, o, do, lse, delta, dqkv, B, T, H)
torch.cuda.synchronize()
_C.lib().abcgpt_debug_attn_trace(0)
for name, base in (("dq", 0), ("dkv", 512)):
    t = tr2[base:base + 512].view(64, 8).cpu()
    print(name, "own step: item n | epilogues  wait_S  compute  store+arrive | total | since previous step start")
    prev = None
    for j in range(64):
        r = t[j]
        if r[4] == 0:
            break
        print(j, int(r[5]), int(r[6]), "|", (r[1] - r[0]).item(), (r[2] - r[1]).item(), (r[3] - r[2]).item(), (r[4] - r[3]).item(), "|",
              (r[4] - r[0]).item(), "|", (r[0] - prev).item() if prev is not None else 0)
        prev = r[0]

m = tr2[1024:1024 + 256].view(64, 4).cpu()
print("dq issuer step: wait_dS  acc_issue  score_issue(incl. kv wait) | since previous")
for j in range(40):
    r = m[j]
    print(j, (r[1] - r[0]).item(), (r[2] - r[1]).item(), (r[3] - r[2]).item(), "|", (r[0] - m[j - 1][0]).item() if j else 0)
