"""The bar to beat on the same box: the UNMODIFIED reference nanoGPT/model.py (vendored copy under
baseline/_ref, tools/vendor_reference.py) under the installed PyTorch on one B200 — bench.py-style loop
(nanoGPT/bench.py:98-117: autocast bf16, model(X, Y), zero_grad, backward, optimizer.step) with fused torch AdamW, eager and
torch.compile, synthetic tokens of the ABC vocabulary, CUDA-event timed, without the per-step loss.item() of the original.

    python tools/bench_torch_ref.py [--workload cfg3|cfg2] [--modes eager,compile] [--steps 20] [--warmup 10] [--clip]

Prints one JSON line per mode (tokens/s, ms/step, MFU against 2.25 PF nominal, clocks).  This is the PyTorch/cuBLAS/SDPA
stack, i.e. library code — a reported comparison, never part of the product path.
"""
from __future__ import annotations

import argparse
import contextlib
import io
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from bench import WORKLOADS, ClockSampler, flops_per_token  # noqa: E402
from tools.vendor_reference import import_reference_model  # noqa: E402


def run(mode, wl, batch, steps, warmup, clip):
    ref = import_reference_model()
    if ref is None:
        return {"mode": mode, "unavailable": "baseline/_ref/nanoGPT/model.py is absent (run tools/vendor_reference.py in the build container)"}
    GPT, GPTConfig = ref
    cfg = wl["cfg"]
    dev = "cuda"
    torch.manual_seed(1337)
    torch.backends.cuda.matmul.allow_tf32 = True
    torch.backends.cudnn.allow_tf32 = True
    with contextlib.redirect_stdout(io.StringIO()):
        model = GPT(GPTConfig(**cfg)).to(dev)
        opt = model.configure_optimizers(weight_decay=0.1, learning_rate=wl["lr"], betas=wl["betas"], device_type="cuda")
    net = torch.compile(model) if mode == "compile" else model
    B, T, V = batch, cfg["block_size"], cfg["vocab_size"]
    g = torch.Generator().manual_seed(1234)
    x = torch.randint(V, (B, T), generator=g).to(dev)
    y = torch.randint(V, (B, T), generator=g).to(dev)
    ctx = torch.amp.autocast(device_type="cuda", dtype=torch.bfloat16)

    def step():
        with ctx:
            _, loss = net(x, y)
        opt.zero_grad(set_to_none=True)
        loss.backward()
        if clip:
            torch.nn.utils.clip_grad_norm_(model.parameters(), 1.0)
        opt.step()
        return loss

    for _ in range(warmup):
        step()
    torch.cuda.synchronize()
    sampler = ClockSampler(0)
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        loss = step()
    e1.record()
    torch.cuda.synchronize()
    clocks = sampler.stop()
    ms = e0.elapsed_time(e1) / steps
    tps = B * T / (ms / 1e3)
    fpt = flops_per_token(cfg)
    return {"impl": "torch-reference", "mode": mode, "workload": wl["name"], "batch": B, "seq_len": T, "tokens_per_s": tps,
            "ms_per_step": ms, "mfu_of_nominal_2250": tps * fpt / 2.25e15, "loss_after": float(loss), "clip": bool(clip),
            "torch": torch.__version__, "peak_mem_gb": torch.cuda.max_memory_allocated() / 2 ** 30, "clocks": clocks,
            "steps": steps, "warmup": warmup}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="cfg3", choices=sorted(WORKLOADS))
    ap.add_argument("--modes", default="eager,compile")
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--batch", type=int, default=0)
    ap.add_argument("--clip", action="store_true", help="add the stock clip_grad_norm_(1.0) of train.py:350-352 to every step")
    args = ap.parse_args()
    wl = WORKLOADS[args.workload]
    for mode in args.modes.split(","):
        try:
            out = run(mode, wl, args.batch or wl["batch"], args.steps, args.warmup, args.clip)
        except Exception as e:  # torch.compile needs a working inductor/triton tool chain on the box: report, do not die
            out = {"impl": "torch-reference", "mode": mode, "workload": wl["name"], "error": f"{type(e).__name__}: {e}"[:400]}
        print(json.dumps(out), flush=True)
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
