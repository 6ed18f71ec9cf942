"""The drop-in claim, tested literally: the reference's own loops with ONLY the import swapped.

  * nanoGPT/bench.py:98-117 (simple benchmarking loop: autocast ctx, model(X, Y), zero_grad, backward, optimizer.step,
    loss.item()) and
  * nanoGPT/train.py:335-357 (gradient accumulation, GradScaler(enabled=False), scaler.unscale_, the STOCK
    torch.nn.utils.clip_grad_norm_(model.parameters(), grad_clip), scaler.step(optimizer), scaler.update(), zero_grad)
are transcribed below statement for statement; `from model import GPTConfig, GPT` becomes
`from ai_music_generation_b200 import GPTConfig, GPT`.  Results are held to the CPU oracle running the same optimizer steps
(bf16-autocast emulation): loss per step |d| <= 5e-3 on step 0 and 2e-2 later (bf16 trajectories drift), gradient norm 3 %.
The data-parallel variant of the train.py loop (our DDP class in place of torch's, everything else stock) runs in
tests/_ddp_gpu_worker.py.  Also here: the module's guard against a second grad-enabled forward before backward, staleness
detection of the bf16 weight shadow, checkpoint resume (train.py:173-216) and rank-sharded estimate_loss (train.py:231-244).
"""
import json
import os
import subprocess
import sys
from contextlib import nullcontext

import numpy as np
import pytest
import torch

from oracle import nanogpt_oracle as O

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CFG = dict(block_size=64, vocab_size=95, n_layer=2, n_head=2, n_embd=128, dropout=0.0, bias=False)


def _model(device, seed=1):
    from ai_music_generation_b200 import GPTConfig, GPT   # <- the only line that differs from the reference scripts
    cfg = O.OracleConfig(**CFG)
    sd = O.synthetic_state(cfg, seed=seed)
    model = GPT(GPTConfig(**CFG))
    model.load_state_dict({**sd, "lm_head.weight": sd["transformer.wte.weight"]})
    model.to(device)
    return model, cfg, sd


def test_reference_bench_loop_verbatim(cuda_device):
    device = str(cuda_device)
    model, cfg, sd = _model(device)
    batches = [O.synthetic_tokens(cfg, 4, 64, seed=s) for s in range(6)]
    it = iter(batches)

    def get_batch(split):
        x, y = next(it)
        return x.pin_memory().to(device, non_blocking=True), y.pin_memory().to(device, non_blocking=True)

    # ---- nanoGPT/bench.py:26-29,60 ------------------------------------------------------------------------------------
    device_type = 'cuda' if 'cuda' in device else 'cpu'
    ptdtype = torch.bfloat16
    ctx = nullcontext() if device_type == 'cpu' else torch.amp.autocast(device_type=device_type, dtype=ptdtype)
    optimizer = model.configure_optimizers(weight_decay=1e-2, learning_rate=1e-3, betas=(0.9, 0.95), device_type=device_type)
    # ---- nanoGPT/bench.py:98-112 (one stage of 5 steps) ---------------------------------------------------------------
    losses = []
    torch.cuda.synchronize()
    X, Y = get_batch('train')
    for k in range(5):
        with ctx:
            logits, loss = model(X, Y)
        X, Y = get_batch('train')
        optimizer.zero_grad(set_to_none=True)
        loss.backward()
        optimizer.step()
        lossf = loss.item()
        losses.append(lossf)
    torch.cuda.synchronize()
    mfu = model.estimate_mfu(4 * 1 * 5, 1.0)
    assert mfu > 0
    # ---- the same five optimizer steps on the oracle (no clipping in bench.py) ------------------------------------------
    ref = O.train_steps({k: v.clone() for k, v in sd.items()}, cfg, batches[:5], lr=1e-3, betas=(0.9, 0.95), weight_decay=1e-2,
                        grad_clip=0.0, bf16=True)
    for k, (got, (want, _)) in enumerate(zip(losses, ref)):
        assert abs(got - want) <= (5e-3 if k == 0 else 2e-2), (k, got, want)


def test_reference_train_loop_verbatim_with_stock_clip(cuda_device):
    device = str(cuda_device)
    model, cfg, sd = _model(device)
    gradient_accumulation_steps, grad_clip, ddp = 2, 1.0, False
    micro = [O.synthetic_tokens(cfg, 2, 64, seed=s) for s in range(9)]
    it = iter(micro)

    def get_batch(split):
        x, y = next(it)
        return x.pin_memory().to(device, non_blocking=True), y.pin_memory().to(device, non_blocking=True)

    device_type = 'cuda'
    ctx = torch.amp.autocast(device_type=device_type, dtype=torch.bfloat16)                      # train.py:116
    scaler = torch.amp.GradScaler("cuda", enabled=False)                                         # train.py:211 (dtype != float16)
    optimizer = model.configure_optimizers(0.1, 1e-3, (0.9, 0.95), device_type)                  # train.py:214
    norms, losses = [], []
    X, Y = get_batch("train")                                                                    # train.py:276
    for iter_num in range(4):
        # ---- nanoGPT/train.py:335-357 -----------------------------------------------------------------------------------
        for micro_step in range(gradient_accumulation_steps):
            if ddp:
                model.require_backward_grad_sync = micro_step == gradient_accumulation_steps - 1
            with ctx:
                logits, loss = model(X, Y)
                loss = loss / gradient_accumulation_steps
            X, Y = get_batch("train")
            scaler.scale(loss).backward()
        if grad_clip != 0.0:
            scaler.unscale_(optimizer)
            norms.append(torch.nn.utils.clip_grad_norm_(model.parameters(), grad_clip).item())
        scaler.step(optimizer)
        scaler.update()
        optimizer.zero_grad(set_to_none=True)
        # ---- train.py:366 ---------------------------------------------------------------------------------------------
        losses.append(loss.item() * gradient_accumulation_steps)
    # oracle: every optimizer step sees the concatenation of its two micro-batches (equal sizes, no ignored targets)
    steps = [(torch.cat((micro[2 * i][0], micro[2 * i + 1][0])), torch.cat((micro[2 * i][1], micro[2 * i + 1][1]))) for i in range(4)]
    ref_sd = {k: v.clone() for k, v in sd.items()}
    ref = O.train_steps(ref_sd, cfg, steps, lr=1e-3, betas=(0.9, 0.95), weight_decay=0.1, grad_clip=1.0, bf16=True)
    for k, (norm, (_, want_norm)) in enumerate(zip(norms, ref)):
        assert norm == pytest.approx(want_norm, rel=3e-2), (k, norm, want_norm)
    # `loss` of the loop is the LAST micro-step's: compare it with the oracle's loss on that micro-batch at the step's weights
    chk = {k: v.clone() for k, v in sd.items()}
    state = {}
    for i in range(4):
        lm, _, _ = O.loss_and_grads(chk, cfg, micro[2 * i + 1][0], micro[2 * i + 1][1], bf16=True)
        assert abs(losses[i] - lm.item()) <= (5e-3 if i == 0 else 2e-2), (i, losses[i], lm.item())
        _, _, grads = O.loss_and_grads(chk, cfg, steps[i][0], steps[i][1], bf16=True)
        c = O.clip_coef(O.grad_norm(grads), 1.0)
        O.adamw_step(chk, {k: v * c for k, v in grads.items()}, state, lr=1e-3, betas=(0.9, 0.95), weight_decay=0.1, step=i + 1)
    # four early AdamW steps move every weight by ~4 lr whatever the gradient's magnitude (update ~ lr * sign(g)), so a bf16
    # sign flip of a tiny gradient element is a full-size difference in that element: compare the UPDATES, not the weights
    named = dict(model.named_parameters())
    for n, want in chk.items():
        got = named[n if n in named else "lm_head.weight"].detach().float().cpu()
        du, dw = (got - sd[n]).flatten(), (want - sd[n]).flatten()
        cos = torch.dot(du, dw) / (du.norm() * dw.norm())
        assert cos.item() >= 0.97, (n, cos.item())
        assert du.norm().item() == pytest.approx(dw.norm().item(), rel=5e-2), n


def test_stock_clip_equals_fused_clip(cuda_device):
    """torch.nn.utils.clip_grad_norm_(model.parameters(), c) + step == model.clip_grad_norm_(c) + step."""
    outs = []
    for fused in (False, True):
        model, cfg, _ = _model(cuda_device)
        opt = model.configure_optimizers(0.1, 1e-3, (0.9, 0.95), "cuda")
        for s in range(3):
            x, y = O.synthetic_tokens(cfg, 4, 64, seed=s)
            _, loss = model(x.to(cuda_device), y.to(cuda_device))
            loss.backward()
            n = model.clip_grad_norm_(0.5) if fused else torch.nn.utils.clip_grad_norm_(model.parameters(), 0.5)
            opt.step()
            opt.zero_grad(set_to_none=True)
        outs.append((n.item(), model._arena["flat"].clone()))
    # not bitwise: the wgrad GEMMs reduce with fp32 atomics, so two runs of the SAME path already differ in the last bits
    # and three optimizer steps amplify that (measured 2e-4 on the third-step norm)
    assert outs[0][0] == pytest.approx(outs[1][0], rel=2e-3)
    assert ((outs[0][1] - outs[1][1]).norm() / outs[1][1].norm()).item() <= 2e-3


def test_second_forward_before_backward_raises_and_losses_are_values(cuda_device):
    model, cfg, _ = _model(cuda_device)
    model.train()
    x1, y1 = (t.to(cuda_device) for t in O.synthetic_tokens(cfg, 4, 64, seed=0))
    x2, y2 = (t.to(cuda_device) for t in O.synthetic_tokens(cfg, 4, 64, seed=1))
    _, l1 = model(x1, y1)
    v1 = l1.item()
    _, l2 = model(x2, y2)
    assert l1.item() == v1 and l2.item() != v1          # every loss is its own tensor (losses.append(loss) keeps values)
    with pytest.raises(RuntimeError, match="overwritten by a later forward"):
        (l1 + l2).backward()
    # the supported orders still work: forward/backward pairs, and an eval forward in between
    _, l1 = model(x1, y1)
    with torch.no_grad():
        model(x2, y2)
    l1.backward()
    assert model._arena["params"][0].grad is not None


def test_weight_edits_outside_the_optimizer_refresh_the_bf16_shadow(cuda_device):
    model, cfg, sd = _model(cuda_device)
    model.eval()
    x, y = (t.to(cuda_device) for t in O.synthetic_tokens(cfg, 4, 64, seed=0))
    _, l0 = model(x, y)
    with torch.no_grad():
        model.transformer.wte.weight.mul_(1.5)            # in-place edit through the parameter: detected by its version
    _, l1 = model(x, y)
    assert abs(l1.item() - l0.item()) > 1e-3
    sd2 = {k: v.clone() for k, v in sd.items()}
    sd2["transformer.wte.weight"] = sd2["transformer.wte.weight"] * 1.5
    ref, _, _ = O.loss_and_grads(sd2, cfg, x.cpu(), y.cpu(), bf16=True)
    assert abs(l1.item() - ref.item()) <= 2e-3
    model.transformer.wte.weight.data.div_(1.5)           # .data writes carry no version bump: the documented escape hatch
    model.mark_weights_dirty()
    _, l2 = model(x, y)
    assert abs(l2.item() - l0.item()) <= 1e-4


def _write_corpus(d, n=6000):
    tune = ("X:1\nT:Test\nM:4/4\nK:D\n|:A2 FA DAFA|B2 GB DBGB|A2 FA DAFA|gece d2 d2:|$\n" * 200)[:n]
    chars = sorted(set(tune))
    stoi = {c: i for i, c in enumerate(chars)}
    ids = np.array([stoi[c] for c in tune], dtype=np.uint16)
    os.makedirs(d, exist_ok=True)
    ids[: int(0.9 * n)].tofile(os.path.join(d, "train.bin"))
    ids[int(0.9 * n):].tofile(os.path.join(d, "val.bin"))
    import pickle
    with open(os.path.join(d, "meta.pkl"), "wb") as f:
        pickle.dump({"vocab_size": len(chars), "itos": {i: c for c, i in stoi.items()}, "stoi": stoi}, f)
    return len(chars)


def test_init_from_resume_continues_the_run(tmp_path, cuda_device):
    """train.py --init_from=resume (reference train.py:173-216): model, optimizer moments, iter_num and best_val_loss come back
    from ckpt.pt; a run of 20 iterations interrupted at 10 and resumed reaches the same state as an uninterrupted run would
    from that checkpoint (same loss at the resume point, optimizer step count continues)."""
    _write_corpus(tmp_path / "data" / "toy")
    env = dict(os.environ, PYTHONPATH=ROOT)
    common = ["--dataset=toy", "--out_dir=out-resume", "--batch_size=8", "--block_size=64", "--n_layer=2", "--n_head=2",
              "--n_embd=128", "--gradient_accumulation_steps=1", "--eval_iters=4", "--log_interval=5",
              "--learning_rate=1e-3", "--warmup_iters=2", "--lr_decay_iters=40", "--min_lr=1e-4", "--dropout=0.0"]
    r = subprocess.run([sys.executable, os.path.join(ROOT, "train.py"), *common, "--max_iters=10", "--eval_interval=10"],
                       cwd=tmp_path, env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    ck = torch.load(tmp_path / "out-resume" / "ckpt.pt", map_location="cpu")
    assert ck["iter_num"] == 10 and "activation" not in ck["model_args"]      # reference-loadable model_args
    assert set(ck["model_args"]) == {"n_layer", "n_head", "n_embd", "block_size", "bias", "vocab_size", "dropout"}
    st = ck["optimizer"]["state"]
    assert float(st[0]["step"]) == 10.0 and st[0]["exp_avg"].abs().sum() > 0
    val_at_10 = float(ck["best_val_loss"])
    r2 = subprocess.run([sys.executable, os.path.join(ROOT, "train.py"), *common, "--init_from=resume", "--max_iters=20",
                         "--eval_interval=10"], cwd=tmp_path, env=env, capture_output=True, text=True, timeout=600)
    assert r2.returncode == 0, r2.stdout[-2000:] + r2.stderr[-2000:]
    assert "Resuming training from out-resume" in r2.stdout
    recs = [json.loads(line) for line in open(tmp_path / "out-resume" / "losses.jsonl")]
    steps = [x["step"] for x in recs]
    assert steps == [0, 10, 10, 20], steps                     # the resumed run re-evaluates at its first iteration (10)
    # same weights => the resumed run's first evaluation reproduces the checkpoint's loss level (different eval batches)
    assert abs(recs[2]["val_loss"] - val_at_10) <= 0.15 * max(1.0, val_at_10)
    assert recs[3]["val_loss"] < recs[2]["val_loss"]            # and it keeps learning
    ck2 = torch.load(tmp_path / "out-resume" / "ckpt.pt", map_location="cpu")
    assert ck2["iter_num"] == 20 and float(ck2["optimizer"]["state"][0]["step"]) == 20.0


def test_estimate_loss_matches_a_plain_loop(cuda_device):
    """evalloop.estimate_loss (world 1) == the reference's loop (train.py:231-244): mean of eval_iters losses per split."""
    from ai_music_generation_b200 import evalloop
    model, cfg, _ = _model(cuda_device)
    model.train()
    batches = {"train": [O.synthetic_tokens(cfg, 4, 64, seed=s) for s in range(5)],
               "val": [O.synthetic_tokens(cfg, 4, 64, seed=10 + s) for s in range(5)]}
    cursor = {"train": 0, "val": 0}

    def get_batch(split):
        x, y = batches[split][cursor[split] % 5]
        cursor[split] += 1
        return x.to(cuda_device), y.to(cuda_device)

    out = evalloop.estimate_loss(model, get_batch, 5, cuda_device)
    assert model.training                                       # train mode restored (train.py:243)
    for split in ("train", "val"):
        want = []
        model.eval()
        with torch.no_grad():
            for x, y in batches[split]:
                want.append(model(x.to(cuda_device), y.to(cuda_device))[1].item())
        model.train()
        assert out[split].item() == pytest.approx(sum(want) / 5, abs=1e-6)
