"""Training driver with the call surface of the reference's nanoGPT/train.py, on the B200 kernels.

    python train.py config/irishman_char.py --batch_size=32
    torchrun --standalone --nproc_per_node=8 train.py config/train_abc_gpt2_small.py

Same settings (reference train.py:37-78), same override mechanism (configurator), same on-disk contracts: flat
uint16 `data/<dataset>/{train,val}.bin` + `meta.pkl` (train.py:122-158), `out_dir/ckpt.pt` with keys model / optimizer /
model_args / iter_num / best_val_loss / config (train.py:316-328), `out_dir/losses.jsonl` (train.py:269-273,307-314), cosine
schedule with warm-up (train.py:248-259), gradient accumulation with the all-reduce only on the last micro-step
(train.py:335-348), clip + AdamW (train.py:350-357), MFU print-out (train.py:360-370).  Differences: the model, optimizer
and DDP wrapper come from ai_music_generation_b200 (fused sm_100a kernels), `compile` is accepted and ignored, `device`
must be a CUDA device.
"""
from __future__ import annotations

import json
import math
import os
import pickle
import sys
import time
from datetime import datetime

import numpy as np
import torch
import torch.distributed as dist

from ai_music_generation_b200 import DDP, GPT, DeviceTokenStream, GPTConfig
from ai_music_generation_b200 import evalloop
from ai_music_generation_b200.data import check_vocab
from configurator import load_settings

DEFAULTS = dict(
    out_dir="out", eval_interval=2000, log_interval=1, eval_iters=200, eval_only=False, always_save_checkpoint=True,
    init_from="scratch", wandb_log=False, wandb_project="owt", wandb_run_name="gpt2", dataset="openwebtext",
    gradient_accumulation_steps=5 * 8, batch_size=12, block_size=1024, n_layer=12, n_head=12, n_embd=768, dropout=0.0,
    bias=False, learning_rate=6e-4, max_iters=600000, weight_decay=1e-1, beta1=0.9, beta2=0.95, grad_clip=1.0,
    decay_lr=True, warmup_iters=2000, lr_decay_iters=600000, min_lr=6e-5, backend="nccl", device="cuda",
    dtype="bfloat16", compile=False,
    device_loader=True,  # new knob: keep train.bin / val.bin in HBM and gather batches with one kernel (same batches as the host loader)
    sharded_eval=True,   # new knob: under DDP every rank evaluates eval_iters / world batches and the sums are all-reduced
                         # (the reference evaluates 2 x eval_iters batches on rank 0 while the other ranks wait, train.py:290-291)
    activation="gelu",   # new knob: "gelu_tanh" = HF gelu_new (TunesFormer-shaped decoders); the reference is exact-erf "gelu"
)


class TokenStream:
    """`get_batch` of the reference (train.py:122-144): random windows of a memory-mapped token file, x and the shifted y
    as int64, pinned and copied asynchronously."""

    def __init__(self, data_dir, block_size, batch_size, device, wide_tokens=False, vocab_size=None):
        self.dir, self.T, self.B, self.device = data_dir, block_size, batch_size, device
        self.dtype = np.uint32 if wide_tokens else np.uint16
        for name in ("train.bin", "val.bin"):  # one pass per file at start-up (see data.check_vocab)
            path = os.path.join(data_dir, name)
            if os.path.exists(path):
                check_vocab(np.memmap(path, dtype=self.dtype, mode="r"), vocab_size, path)

    def get(self, split):
        data = np.memmap(os.path.join(self.dir, "train.bin" if split == "train" else "val.bin"), dtype=self.dtype, mode="r")
        starts = torch.randint(len(data) - self.T, (self.B,)).tolist()
        x = torch.from_numpy(np.stack([data[i:i + self.T] for i in starts]).astype(np.int64))
        y = torch.from_numpy(np.stack([data[i + 1:i + 1 + self.T] for i in starts]).astype(np.int64))
        return (x.pin_memory().to(self.device, non_blocking=True), y.pin_memory().to(self.device, non_blocking=True))


def lr_at(it, s):
    if it < s["warmup_iters"]:
        return s["learning_rate"] * (it + 1) / (s["warmup_iters"] + 1)
    if it > s["lr_decay_iters"]:
        return s["min_lr"]
    ratio = (it - s["warmup_iters"]) / (s["lr_decay_iters"] - s["warmup_iters"])
    return s["min_lr"] + 0.5 * (1.0 + math.cos(math.pi * ratio)) * (s["learning_rate"] - s["min_lr"])


def main():
    s = load_settings(DEFAULTS, sys.argv[1:])
    config = {k: v for k, v in s.items() if isinstance(v, (int, float, bool, str))}
    ddp = int(os.environ.get("RANK", -1)) != -1
    if ddp:
        dist.init_process_group(backend=s["backend"])
        rank, local_rank, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
        device = f"cuda:{local_rank}"
        torch.cuda.set_device(device)
        assert s["gradient_accumulation_steps"] % world == 0
        accum = s["gradient_accumulation_steps"] // world
    else:
        rank, world, device, accum = 0, 1, s["device"], s["gradient_accumulation_steps"]
    master = rank == 0
    if "cuda" not in device:
        raise SystemExit("train.py: the sm_100a path needs a CUDA device (the reference's --device=cpu run is the CPU baseline)")
    print(f"tokens per iteration will be: {accum * world * s['batch_size'] * s['block_size']:,}")
    if master:
        os.makedirs(s["out_dir"], exist_ok=True)
    torch.manual_seed(1337 + rank)

    data_dir = os.path.join("data", s["dataset"])
    vocab = None
    meta_path = os.path.join(data_dir, "meta.pkl")
    if os.path.exists(meta_path):
        with open(meta_path, "rb") as f:
            vocab = pickle.load(f)["vocab_size"]
        print(f"found vocab_size = {vocab} (inside {meta_path})")
    Loader = DeviceTokenStream if s["device_loader"] else TokenStream
    stream = Loader(data_dir, s["block_size"], s["batch_size"], device,
                    wide_tokens=s["out_dir"] == "out-irishman-whitespace",  # the reference's uint32 special case
                    vocab_size=vocab)

    # `model_args` goes into ckpt.pt and the reference's sample.py does GPTConfig(**checkpoint['model_args'])
    # (nanoGPT/sample.py:59): the extra `activation` key is written only when it differs from the reference's GELU, so every
    # checkpoint of a reference-shaped model loads in the reference unchanged
    model_args = dict(n_layer=s["n_layer"], n_head=s["n_head"], n_embd=s["n_embd"], block_size=s["block_size"],
                      bias=s["bias"], vocab_size=None, dropout=s["dropout"])
    if s["activation"] != "gelu":
        model_args["activation"] = s["activation"]
    iter_num, best_val = 0, 1e9
    checkpoint = None
    if s["init_from"] == "scratch":
        print("Initializing a new model from scratch")
        model_args["vocab_size"] = vocab if vocab is not None else 50304
        model = GPT(GPTConfig(**model_args))
    elif s["init_from"] == "resume":
        print(f"Resuming training from {s['out_dir']}")
        checkpoint = torch.load(os.path.join(s["out_dir"], "ckpt.pt"), map_location="cpu")
        for k in ("n_layer", "n_head", "n_embd", "block_size", "bias", "vocab_size"):
            model_args[k] = checkpoint["model_args"][k]
        if checkpoint["model_args"].get("activation", "gelu") != "gelu":
            model_args["activation"] = checkpoint["model_args"]["activation"]
        else:
            model_args.pop("activation", None)
        model = GPT(GPTConfig(**model_args))
        sd = {k.removeprefix("_orig_mod."): v for k, v in checkpoint["model"].items()}
        model.load_state_dict(sd)
        iter_num, best_val = checkpoint["iter_num"], checkpoint["best_val_loss"]
    elif s["init_from"].startswith("gpt2"):
        # nanoGPT/train.py:197-205: OpenAI GPT-2 weights through transformers (needs the HF cache or network access)
        print(f"Initializing from OpenAI GPT-2 weights: {s['init_from']}")
        model = GPT.from_pretrained(s["init_from"], dict(dropout=s["dropout"]))
        for k in ("n_layer", "n_head", "n_embd", "block_size", "bias", "vocab_size"):
            model_args[k] = getattr(model.config, k)
    else:
        raise SystemExit(f"unknown init_from {s['init_from']!r} (scratch | resume | gpt2*)")
    if s["block_size"] < model.config.block_size:
        model.crop_block_size(s["block_size"])
        model_args["block_size"] = s["block_size"]
    model.to(device)
    optimizer = model.configure_optimizers(s["weight_decay"], s["learning_rate"], (s["beta1"], s["beta2"]), "cuda")
    if checkpoint is not None:
        optimizer.load_state_dict(checkpoint["optimizer"])
    checkpoint = None
    raw_model = model
    if ddp:
        model = DDP(model, device_ids=[local_rank])

    shard_eval = ddp and s["sharded_eval"]

    def estimate_loss():
        return evalloop.estimate_loss(raw_model, stream.get, s["eval_iters"], device, shard=shard_eval)

    log_path = os.path.join(s["out_dir"], "losses.jsonl")
    if master and not os.path.exists(log_path):
        open(log_path, "w").close()

    X, Y = stream.get("train")
    t0 = time.time()
    local_iter, running_mfu = 0, -1.0
    while True:
        lr = lr_at(iter_num, s) if s["decay_lr"] else s["learning_rate"]
        for group in optimizer.param_groups:
            group["lr"] = lr
        if iter_num % s["eval_interval"] == 0 and (master or shard_eval):
            losses = estimate_loss()
        if iter_num % s["eval_interval"] == 0 and master:
            print(f"[{datetime.now().strftime('%H:%M:%S')}] step {iter_num}: train loss {losses['train']:.4f}, "
                  f"val loss {losses['val']:.4f}")
            with open(log_path, "a") as f:
                f.write(json.dumps({"step": iter_num, "train_loss": losses["train"].item(), "val_loss": losses["val"].item()}) + "\n")
            if losses["val"] < best_val or s["always_save_checkpoint"]:
                best_val = losses["val"]
                if iter_num > 0:
                    checkpoint = {"model": raw_model.state_dict(), "optimizer": optimizer.state_dict(), "model_args": model_args,
                                  "iter_num": iter_num, "best_val_loss": best_val, "config": config}
                    print(f"saving checkpoint to {s['out_dir']}")
                    torch.save(checkpoint, os.path.join(s["out_dir"], "ckpt.pt"))
            torch.save(checkpoint, os.path.join(s["out_dir"], "last_iter_ckpt.pt"))  # same quirk as the reference (:329)
        if iter_num == 0 and s["eval_only"]:
            break

        for micro in range(accum):
            if ddp:
                model.require_backward_grad_sync = micro == accum - 1
            _, loss = model(X, Y)
            loss = loss / accum
            X, Y = stream.get("train")  # host work overlaps the asynchronous forward
            loss.backward()
        if s["grad_clip"] != 0.0:
            raw_model.clip_grad_norm_(s["grad_clip"])  # fused form of torch.nn.utils.clip_grad_norm_(model.parameters(), …)
        optimizer.step()
        optimizer.zero_grad(set_to_none=True)

        t1 = time.time()
        dt, t0 = t1 - t0, t1
        if iter_num % s["log_interval"] == 0 and master:
            lossf = loss.item() * accum  # device sync, as in the reference
            if local_iter >= 5:
                mfu = raw_model.estimate_mfu(s["batch_size"] * accum, dt, flops_promised=2.25e15)
                running_mfu = mfu if running_mfu == -1.0 else 0.9 * running_mfu + 0.1 * mfu
            print(f"iter {iter_num}: loss {lossf:.4f}, time {dt * 1000:.2f}ms, "
                  f"{s['batch_size'] * accum * s['block_size'] * world / dt:,.0f} tok/s, mfu(B200 2.25PF) {running_mfu * 100:.2f}%")
        iter_num += 1
        local_iter += 1
        if iter_num > s["max_iters"]:
            break
    if ddp:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
