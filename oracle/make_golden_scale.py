"""BASELINE-scale golden vectors from the UNMODIFIED reference (run in the build container only; ~45 min of CPU).

    python oracle/make_golden_scale.py [step] [cfg5]      # needs /root/reference/nanoGPT/model.py

Two fixtures at the GPT-2-small shape of BASELINE.json configs[2] / configs[4] (12L / 12H / 768d, block 1024, V = 95):

  tests/golden/nanogpt_gpt2s_step.json      one training step at B = 4, T = 1024 on the closed-form weights / tokens of
      oracle/nanogpt_oracle.py: loss, logits slices, per-tensor gradient norms AND eight sampled gradient elements per tensor,
      total norm, parameter norms after clip + AdamW; the same step under CPU bf16 autocast gives the tolerance yardsticks
      (loss gap, logits max / mean gap, per-tensor gradient relative L2 gap of the reference against itself).

  tests/golden/nanogpt_cfg5_generate.json + nanogpt_cfg5_weights.npz      sample.py's batched greedy generation
      (model.generate(idx[4, 1], 1024, temperature=1.0, top_k=1), nanoGPT/model.py:305-330; the last token slides the
      window) on SHARPENED weights, so that "identical token ids" is a meaningful check (on unsharpened weights the
      reference's own bf16 run leaves its fp32 run within a few tokens).  Sharpening = the reference trained
      on a cyclic tune of 64 distinct symbols; only a small parameter subset is trained so that the EXACT weights can be
      committed: wte (tied head), every LayerNorm weight and a rank-16 update U V of the last block's mlp.c_proj (~170 k
      numbers, stored in the .npz; everything else is the closed-form state).  Also stored: the reference's top-2 margin at
      every generated position (teacher-forced forward over the generated tunes) and its last-position logits at probe
      positions, which the GPU test compares with the KV-cache decode path's logits.

The fixtures hold only outputs and the trained subset, never reference source.
"""
from __future__ import annotations

import contextlib
import io
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
REF = "/root/reference/nanoGPT"
GOLDEN = os.path.join(ROOT, "tests", "golden")

from oracle import nanogpt_oracle as O  # noqa: E402

CFG = dict(block_size=1024, vocab_size=95, n_layer=12, n_head=12, n_embd=768, dropout=0.0, bias=False)
STEP = dict(cfg=CFG, batch=4, seqlen=1024, lr=6e-4, betas=(0.9, 0.95), seed=1)
# 580 steps on 2 x 256-token windows, then 20 on one full 1024-token window (all positions seen); cosine decay to lr / 10
SHARPEN = dict(rank=16, steps=600, long_steps=20, lr=1e-2, betas=(0.9, 0.95), batch=2, seqlen=256, period=64, seed=1)
PROMPT_LEN, NEW_TOKENS = 3, 1024   # 3 + 1024 > block_size: the last two tokens come from a slid window (model.py:314)
PROBES = (2, 3, 63, 64, 255, 511, 767, 1022, 1023)   # context positions whose next-token logits are stored


def reference_model():
    sys.path.insert(0, REF)
    from model import GPT, GPTConfig  # the reference, unmodified
    with contextlib.redirect_stdout(io.StringIO()):
        return GPT(GPTConfig(**CFG))


def grad_sample_index(numel):
    """Eight fixed element positions per tensor (spread over the tensor) whose gradient values are pinned."""
    return [(i * 2654435761 + 12345) % numel for i in range(8)]


def make_step():
    cfg = O.OracleConfig(**CFG)
    model = reference_model()
    sd = O.synthetic_state(cfg, seed=STEP["seed"])
    model.load_state_dict({**sd, "lm_head.weight": sd["transformer.wte.weight"]}, strict=True)
    model.train()
    with contextlib.redirect_stdout(io.StringIO()):
        opt = model.configure_optimizers(0.1, STEP["lr"], STEP["betas"], "cpu")
    names = O.param_names(cfg)
    named = dict(model.named_parameters())
    x, y = O.synthetic_tokens(cfg, STEP["batch"], STEP["seqlen"], seed=0)
    t0 = time.time()
    logits, loss = model(x, y)
    opt.zero_grad(set_to_none=True)
    loss.backward()
    print(f"fp32 step {time.time() - t0:.1f}s loss {loss.item():.6f}", flush=True)
    g32 = {n: named[n].grad.detach().clone() for n in names}
    rec = {"loss": loss.item(), "logits_slice": logits[0, :4, :8].detach().flatten().tolist(),
           "logits_last_slice": logits[:, -1, :8].detach().flatten().tolist(),
           "logits_absmean": logits.detach().abs().mean().item(), "logits_checksum": logits.detach().double().sum().item(),
           "grad_norms": {n: g32[n].norm().item() for n in names},
           "grad_samples": {n: g32[n].flatten()[grad_sample_index(g32[n].numel())].tolist() for n in names}}
    logits32 = logits.detach().clone()
    total = torch.nn.utils.clip_grad_norm_(model.parameters(), 1.0)
    rec["grad_norm_total"] = total.item()
    opt.step()
    rec["param_norms_after"] = {n: named[n].detach().norm().item() for n in names}
    # the reference against itself under bf16 autocast: the yardstick for the bf16 GPU path at this scale
    model.load_state_dict({**sd, "lm_head.weight": sd["transformer.wte.weight"]}, strict=True)
    opt.zero_grad(set_to_none=True)
    t0 = time.time()
    with torch.autocast("cpu", dtype=torch.bfloat16):
        lb, lossb = model(x, y)
    lossb.backward()
    print(f"bf16 autocast step {time.time() - t0:.1f}s loss {lossb.item():.6f}", flush=True)
    d = (lb.float() - logits32).abs()
    rec["bf16_autocast"] = {
        "loss": lossb.item(), "logits_maxdiff": d.max().item(), "logits_meandiff": d.mean().item(),
        "grad_rel_l2": {n: ((named[n].grad - g32[n]).norm() / g32[n].norm()).item() for n in names},
        "grad_norm_total": torch.sqrt(sum((named[n].grad.double() ** 2).sum() for n in names)).item()}
    out = {"spec": {**STEP, "betas": list(STEP["betas"])}, "step": rec, "num_params": model.get_num_params()}
    with open(os.path.join(GOLDEN, "nanogpt_gpt2s_step.json"), "w") as f:
        json.dump(out, f, indent=1)
    print("wrote nanogpt_gpt2s_step.json", flush=True)


def tune(offset, n, period):
    perm = [(17 * i + 3) % 95 for i in range(period)]   # distinct symbols: 17 is coprime to 95
    return torch.tensor([perm[(offset + i) % period] for i in range(n)], dtype=torch.int64)


def make_cfg5():
    cfg = O.OracleConfig(**CFG)
    model = reference_model()
    sd = O.synthetic_state(cfg, seed=SHARPEN["seed"])
    model.load_state_dict({**sd, "lm_head.weight": sd["transformer.wte.weight"]}, strict=True)
    for p in model.parameters():
        p.requires_grad_(False)
    base = {n: p.detach() for n, p in model.named_parameters()}
    L, C, r, T = cfg.n_layer, cfg.n_embd, SHARPEN["rank"], cfg.block_size
    target = f"transformer.h.{L - 1}.mlp.c_proj.weight"
    wte = base["transformer.wte.weight"].clone().requires_grad_(True)
    lns = {n: p.clone().requires_grad_(True) for n, p in base.items() if "ln_" in n}
    U = torch.zeros(C, r, requires_grad=True)
    V = (torch.randn(r, 4 * C, generator=torch.Generator().manual_seed(0)) * 0.02).requires_grad_(True)
    opt = torch.optim.AdamW([wte, U, V, *lns.values()], lr=SHARPEN["lr"], betas=SHARPEN["betas"], weight_decay=0.0)
    import math
    for step in range(SHARPEN["steps"]):
        t0 = time.time()
        long = step >= SHARPEN["steps"] - SHARPEN["long_steps"]
        nb, nt = (1, T) if long else (SHARPEN["batch"], SHARPEN["seqlen"])
        for gp in opt.param_groups:
            gp["lr"] = SHARPEN["lr"] * (0.1 + 0.9 * 0.5 * (1 + math.cos(math.pi * step / SHARPEN["steps"])))
        s = torch.stack([tune((7 * step + 13 * b) % SHARPEN["period"], nt + 1, SHARPEN["period"]) for b in range(nb)])
        params = {**base, **lns, "transformer.wte.weight": wte, "lm_head.weight": wte, target: base[target] + U @ V}
        _, loss = torch.func.functional_call(model, params, (s[:, :-1].contiguous(), s[:, 1:].contiguous()))
        opt.zero_grad()
        loss.backward()
        opt.step()
        if step % 20 == 0 or long:
            print(f"sharpen {step}: loss {loss.item():.4f} ({time.time() - t0:.1f}s)", flush=True)
    final_loss = loss.item()
    # commit the trained subset; the low-rank product is formed in a fixed order (rank-1 terms, fp64, then rounded) so that
    # the test reconstructs bit-identical fp32 weights on any machine
    Un, Vn = U.detach().numpy(), V.detach().numpy()
    np.savez_compressed(os.path.join(GOLDEN, "nanogpt_cfg5_weights.npz"), wte=wte.detach().numpy(), U=Un, V=Vn,
                        **{"ln:" + n: p.detach().numpy() for n, p in lns.items()})
    sharp = sharpened_state(sd, os.path.join(GOLDEN, "nanogpt_cfg5_weights.npz"), cfg)
    model.load_state_dict({**sharp, "lm_head.weight": sharp["transformer.wte.weight"]}, strict=True)
    model.eval()
    prompt = torch.stack([tune(o, PROMPT_LEN, SHARPEN["period"]) for o in (0, 5, 21, 40)])
    t0 = time.time()
    with torch.no_grad():
        gen = model.generate(prompt, NEW_TOKENS, temperature=1.0, top_k=1)
    print(f"reference generate: {time.time() - t0:.0f}s", flush=True)
    total = PROMPT_LEN + NEW_TOKENS
    assert gen.shape == (4, total)
    with torch.no_grad():
        # teacher-forced margins over the non-sliding part: logits at context position t given gen[:, :t + 1] are what
        # generate saw when it produced token t + 1 (t + 1 <= block_size)
        logits, _ = model(gen[:, :T].contiguous(), gen[:, 1:T + 1].contiguous())
        top2 = torch.topk(logits, 2, dim=-1).values
        margins = (top2[..., 0] - top2[..., 1])[:, PROMPT_LEN - 1:]
        agree = (logits.argmax(-1) == gen[:, 1:T + 1])[:, PROMPT_LEN - 1:]
        probes = {str(t): logits[:, t, :].tolist() for t in PROBES}
        slid_margins = []
        for i in range(T + 1, total):   # token i was produced from the slid window gen[:, i - T:i]
            ls, _ = model(gen[:, i - T:i].contiguous())
            t2 = torch.topk(ls[:, -1, :], 2, dim=-1).values
            slid_margins.append((t2[:, 0] - t2[:, 1]).tolist())
            agree = agree & (ls[:, -1, :].argmax(-1) == gen[:, i]).all()
            probes[f"slid_{i}"] = ls[:, -1, :].tolist()
    out = {"spec": {"cfg": CFG, "sharpen": {**SHARPEN, "betas": list(SHARPEN["betas"])}, "target": target,
                    "final_sharpen_loss": final_loss},
           "prompt": prompt.tolist(), "new_tokens": NEW_TOKENS, "tokens": gen.tolist(),
           "teacher_forced_agrees": bool(agree.all()), "margin_min": margins.min().item(),
           "margin_min_per_tune": margins.min(dim=1).values.tolist(),
           "margin_slid": slid_margins,
           "probe_logits": probes}
    with open(os.path.join(GOLDEN, "nanogpt_cfg5_generate.json"), "w") as f:
        json.dump(out, f)
    print("wrote nanogpt_cfg5_generate.json: margin_min", out["margin_min"], "teacher-forced agrees", out["teacher_forced_agrees"],
          flush=True)


def sharpened_state(sd, npz_path, cfg):
    """closed-form state + the committed trained subset -> the exact fp32 state the reference generated from."""
    z = np.load(npz_path)
    out = {k: v.clone() for k, v in sd.items()}
    out["transformer.wte.weight"] = torch.from_numpy(z["wte"]).clone()
    for k in z.files:
        if k.startswith("ln:"):
            out[k[3:]] = torch.from_numpy(z[k]).clone()
    U, V = z["U"].astype(np.float64), z["V"].astype(np.float64)
    delta = np.zeros((U.shape[0], V.shape[1]), dtype=np.float64)
    for j in range(U.shape[1]):
        delta += np.outer(U[:, j], V[j])
    name = f"transformer.h.{cfg.n_layer - 1}.mlp.c_proj.weight"
    out[name] = (out[name].double() + torch.from_numpy(delta)).float()
    return out


def main():
    torch.manual_seed(0)
    torch.set_num_threads(int(os.environ.get("GOLDEN_THREADS", "8")))
    os.makedirs(GOLDEN, exist_ok=True)
    what = sys.argv[1:] or ["step", "cfg5"]
    if "step" in what:
        make_step()
    if "cfg5" in what:
        make_cfg5()


if __name__ == "__main__":
    main()
