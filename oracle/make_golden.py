"""Generates tests/golden/*.json from the UNMODIFIED reference (run in the build container only).

    python oracle/make_golden.py            # needs /root/reference/nanoGPT/model.py

The reference ships no golden vectors of its own, so these are produced by importing its
model.py, loading the closed-form synthetic weights / token streams defined in oracle/nanogpt_oracle.py, and
running the reference's own forward / backward / clip_grad_norm_ / AdamW(configure_optimizers) / generate on
CPU in fp32 (and forward under CPU bf16 autocast for tolerance grounding).  The fixtures hold only outputs
(scalars, small slices, per-tensor norms), never reference source.
"""
from __future__ import annotations

import contextlib
import io
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
REF = "/root/reference/nanoGPT"

from oracle import nanogpt_oracle as O  # noqa: E402

CASES = {
    "tiny": dict(cfg=dict(block_size=64, vocab_size=95, n_layer=2, n_head=2, n_embd=128, dropout=0.0, bias=False),
                 batch=4, seqlen=64, lr=1e-3, betas=(0.9, 0.99), steps=3, gen_prompt=8, gen_new=24),
    "tiny_bias": dict(cfg=dict(block_size=64, vocab_size=95, n_layer=2, n_head=2, n_embd=128, dropout=0.0, bias=True),
                      batch=3, seqlen=48, lr=6e-4, betas=(0.9, 0.95), steps=2, gen_prompt=4, gen_new=8),
    "baby": dict(cfg=dict(block_size=256, vocab_size=95, n_layer=6, n_head=6, n_embd=384, dropout=0.0, bias=False),
                 batch=4, seqlen=256, lr=1e-3, betas=(0.9, 0.99), steps=2, gen_prompt=1, gen_new=12),
    "ignore_index": dict(cfg=dict(block_size=32, vocab_size=95, n_layer=1, n_head=1, n_embd=64, dropout=0.0, bias=False),
                         batch=2, seqlen=32, lr=1e-3, betas=(0.9, 0.95), steps=1, gen_prompt=2, gen_new=4,
                         ignore_every=3),
}


def run_case(name, spec):
    sys.path.insert(0, REF)
    from model import GPT, GPTConfig  # the reference, unmodified
    cfg = O.OracleConfig(**spec["cfg"])
    with contextlib.redirect_stdout(io.StringIO()):
        model = GPT(GPTConfig(**spec["cfg"]))
    sd = O.synthetic_state(cfg, seed=1)
    missing = model.load_state_dict({**sd, "lm_head.weight": sd["transformer.wte.weight"]}, strict=True)
    assert not missing.missing_keys and not missing.unexpected_keys
    model.train()
    with contextlib.redirect_stdout(io.StringIO()):
        opt = model.configure_optimizers(0.1, spec["lr"], spec["betas"], "cpu")
    out = {"spec": {**spec, "betas": list(spec["betas"])}, "steps": []}
    names = O.param_names(cfg)
    named = dict(model.named_parameters())
    for step in range(spec["steps"]):
        x, y = O.synthetic_tokens(cfg, spec["batch"], spec["seqlen"], seed=step)
        if spec.get("ignore_every"):
            y = y.clone()
            y.view(-1)[:: spec["ignore_every"]] = -1
        logits, loss = model(x, y)
        opt.zero_grad(set_to_none=True)
        loss.backward()
        rec = {"loss": loss.item()}
        if step == 0:
            rec["logits_slice"] = logits[0, :4, :8].detach().flatten().tolist()
            rec["logits_absmean"] = logits.detach().abs().mean().item()
            rec["logits_checksum"] = logits.detach().double().sum().item()
            with torch.autocast("cpu", dtype=torch.bfloat16):
                lb, lossb = model(x, y)
            rec["bf16_autocast_loss"] = lossb.item()
            rec["bf16_autocast_logits_maxdiff"] = (lb.float() - logits).abs().max().item()
        rec["grad_norms"] = {n: named[n if n in named else "lm_head.weight"].grad.norm().item() for n in names}
        total = torch.nn.utils.clip_grad_norm_(model.parameters(), 1.0)
        rec["grad_norm_total"] = total.item()
        opt.step()
        rec["param_norms_after"] = {n: named[n if n in named else "lm_head.weight"].detach().norm().item() for n in names}
        out["steps"].append(rec)
    model.eval()
    prompt, _ = O.synthetic_tokens(cfg, 2, spec["gen_prompt"], seed=99)
    gen = model.generate(prompt, spec["gen_new"], temperature=1.0, top_k=1)
    out["generate"] = {"prompt": prompt.tolist(), "tokens": gen.tolist()}
    # top-2 margins of the reference at every generated position (positions inside the bf16 tolerance are reported,
    # not failed, by the GPU parity test)
    _, margins = O.generate_greedy({k: v.detach().clone() for k, v in named_state(model, names).items()}, cfg, prompt,
                                   spec["gen_new"], return_margins=True)
    out["generate"]["oracle_margins_min"] = margins.min().item()
    out["num_params"] = model.get_num_params()
    out["mfu_at_1s_per_iter"] = model.estimate_mfu(1, 1.0)
    return out


def named_state(model, names):
    sd = model.state_dict()
    return {n: sd[n] for n in names}


def main():
    torch.manual_seed(0)
    torch.set_num_threads(8)
    os.makedirs(os.path.join(ROOT, "tests", "golden"), exist_ok=True)
    for name, spec in CASES.items():
        res = run_case(name, spec)
        path = os.path.join(ROOT, "tests", "golden", f"nanogpt_{name}.json")
        with open(path, "w") as f:
            json.dump(res, f, indent=1)
        print(name, [s["loss"] for s in res["steps"]], os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
