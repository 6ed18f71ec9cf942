"""Per-CTA timeline of the attention kernels (cfg3 shape): start/end (globaltimer ns), SM id and step count of every CTA.
Prints, per kernel: span, mean CTA duration by step count, per-SM busy fraction and the gap between consecutive CTAs."""
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
do = torch.randn(B * T, C, device="cuda").bfloat16()
dqkv = torch.empty(B * T, 3 * C, device="cuda", dtype=torch.bfloat16)
delta = torch.empty(B, H, T, device="cuda")
for _ in range(3):
    ops.attn_fwd(qkv, o, lse, B, T, H)
    ops.attn_bwd(qkv, o, do, lse, delta, dqkv, B, T, H)
ncta = 1024  # slots per kernel (persistent grids: <= 2 CTAs per SM)
tr = torch.zeros(3 * ncta * 4, device="cuda", dtype=torch.int64)
_C.lib().abcgpt_debug_attn_cta_trace(tr.data_ptr())
ops.attn_fwd(qkv, o, lse, B, T, H)
ops.attn_bwd(qkv, o, do, lse, delta, dqkv, B, T, H)
torch.cuda.synchronize()
_C.lib().abcgpt_debug_attn_cta_trace(0)
t = tr.view(3, ncta, 4).cpu()
for k, name in enumerate(("fwd", "dkv", "dq")):
    r = t[k]
    r = r[r[:, 1] > 0]
    t0, t1, sm, steps = r[:, 0], r[:, 1], r[:, 2], r[:, 3]
    span = (t1.max() - t0.min()).item()
    dur = (t1 - t0).float()
    print(f"== {name}: span {span / 1e3:.1f} us, {len(r)} CTAs, steps/CTA {steps.min().item()}..{steps.max().item()}, ns/step {(dur / steps.float()).mean().item():.0f}, mean CTA {dur.mean().item() / 1e3:.2f} us, sum/SMs {dur.sum().item() / 148 / 1e3:.1f} us")
    gaps, busy = [], []
    for s_id in sorted(set(sm.tolist())):
        m = sm == s_id
        a0, a1 = t0[m], t1[m]
        order = a0.argsort()
        a0, a1 = a0[order], a1[order]
        busy.append((a1 - a0).sum().item() / span)
        if name != "fwd" and len(a0) > 1:
            gaps += (a0[1:] - a1[:-1]).tolist()
    print(f"   per-SM busy fraction (sum of CTA durations / span): mean {sum(busy) / len(busy):.3f}")
    if gaps:
        g = torch.tensor(gaps).float()
        print(f"   gap between consecutive CTAs on one SM: mean {g.mean().item():.0f} ns, median {g.median().item():.0f}, max {g.max().item():.0f}")
    print(f"   first start spread {(t0.sort().values[147] - t0.min()).item()} ns; last-end minus median-end of final CTAs per SM: see busy")
