"""BASELINE config 4 on one B200 (or, under torchrun, data-parallel over N B200s with TunesFormerDDP: weak scaling, one
shard of --tunes tunes per GPU, CUDA-event time = max over ranks): TunesFormer-shaped decoder (patch level 9 layers T = 128 patches, char level 3 layers T = 32,
768-d, bias, tanh GELU) on synthetic bar-patched tunes, one optimizer step = forward + backward + joint clip + AdamW.
Prints one JSON line (characters/s = non-pad and pad character positions of the char-level decoder per second)."""
import argparse
import contextlib
import io
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ai_music_generation_b200 import GPTConfig, TunesFormerDDP, TunesFormerShaped, ops  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--tunes", type=int, default=64)
ap.add_argument("--patches", type=int, default=128)
ap.add_argument("--steps", type=int, default=10)
ap.add_argument("--warmup", type=int, default=3)
ap.add_argument("--breakdown", action="store_true", help="extra instrumented pass: per-kernel-family milliseconds per step")
args = ap.parse_args()
rank, local, world = int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
kw = dict(n_head=12, n_embd=768, dropout=0.0, bias=True, activation="gelu_tanh")
torch.manual_seed(1337)
with contextlib.redirect_stdout(io.StringIO()):
    model = TunesFormerShaped(GPTConfig(block_size=128, vocab_size=1, n_layer=9, **kw),
                              GPTConfig(block_size=32, vocab_size=128, n_layer=3, **kw)).to(dev).train()
    opt = model.configure_optimizers(0.01, 5e-5, (0.9, 0.999))
fwd = TunesFormerDDP(model) if world > 1 else model
g = torch.Generator().manual_seed(1234 + rank)
B, P = args.tunes, args.patches
patches = torch.randint(3, 128, (B, P, 32), generator=g)
lens = torch.randint(8, 33, (B, P), generator=g)
patches[torch.arange(32)[None, None, :] >= lens[..., None]] = 0
patches = patches.to(dev)


def step():
    loss = fwd(patches)
    opt.zero_grad(set_to_none=True)
    loss.backward()
    model.clip_grad_norm_(1.0)
    opt.step()
    return loss


for _ in range(args.warmup):
    step()
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
l0 = ops.LAUNCHES
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(args.steps):
    loss = step()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / args.steps
if world > 1:
    t = torch.tensor([ms], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = t.item()
fam = None
if args.breakdown:
    prof = []
    ops.set_profile(prof)
    step()
    torch.cuda.synchronize()
    ops.set_profile(None)
    fam = {}
    for name, meta, a0, a1 in prof:
        key = name
        if name == "gemm":
            key = {(0, 0): "gemm_fwd", (0, 1): "gemm_dgrad", (1, 1): "gemm_wgrad"}.get((meta[3], meta[4]), "gemm")
        fam[key] = round(fam.get(key, 0.0) + a0.elapsed_time(a1), 3)
chars = world * B * (P - 1) * 32
n_params = sum(p.numel() for p in model.parameters())
if rank == 0:
    print(json.dumps({"n_gpus": world, "workload": f"TunesFormer-shaped 9L(T=128 patches)+3L(T=32 chars) 768d, {B} tunes x {P} patches x 32 chars",
                      "ms_per_step": ms, "chars_per_s": chars / ms * 1e3, "patches_per_s": world * B * P / ms * 1e3,
                      "params": n_params, "gpu_launches_per_step": (ops.LAUNCHES - l0) // args.steps, "loss": loss.item(),
                      **({"kernel_breakdown_ms": fam} if fam else {})}))
if world > 1:
    dist.destroy_process_group()
