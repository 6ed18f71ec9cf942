"""Pins the CPU oracle (oracle/nanogpt_oracle.py) to vectors produced by the unmodified reference
(oracle/make_golden.py -> tests/golden/nanogpt_*.json).  CPU only."""
import json
import math
import os

import pytest
import torch

from oracle import nanogpt_oracle as O

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
CASES = ["tiny", "tiny_bias", "baby", "ignore_index"]


def load(name):
    with open(os.path.join(GOLDEN, f"nanogpt_{name}.json")) as f:
        return json.load(f)


def batch_for(cfg, spec, step):
    x, y = O.synthetic_tokens(cfg, spec["batch"], spec["seqlen"], seed=step)
    if spec.get("ignore_every"):
        y = y.clone()
        y.view(-1)[:: spec["ignore_every"]] = -1
    return x, y


@pytest.mark.parametrize("name", CASES)
def test_oracle_training_steps_match_reference(name):
    torch.set_num_threads(8)
    g = load(name)
    spec = g["spec"]
    cfg = O.OracleConfig(**spec["cfg"])
    sd = O.synthetic_state(cfg, seed=1)
    state = {}
    for step, rec in enumerate(g["steps"]):
        x, y = batch_for(cfg, spec, step)
        loss, logits, grads = O.loss_and_grads(sd, cfg, x, y)
        assert loss.item() == pytest.approx(rec["loss"], rel=2e-6, abs=2e-6)
        if step == 0:
            got = logits[0, :4, :8].flatten()
            assert torch.allclose(got, torch.tensor(rec["logits_slice"]), rtol=1e-4, atol=2e-6)
            assert logits.abs().mean().item() == pytest.approx(rec["logits_absmean"], rel=1e-5)
        for n, ref_norm in rec["grad_norms"].items():
            assert grads[n].norm().item() == pytest.approx(ref_norm, rel=2e-4, abs=1e-7), n
        total = O.grad_norm(grads)
        assert total == pytest.approx(rec["grad_norm_total"], rel=1e-4)
        c = O.clip_coef(total, 1.0)
        grads = {k: v * c for k, v in grads.items()}
        O.adamw_step(sd, grads, state, lr=spec["lr"], betas=tuple(spec["betas"]), weight_decay=0.1, step=step + 1)
        for n, ref_norm in rec["param_norms_after"].items():
            assert sd[n].norm().item() == pytest.approx(ref_norm, rel=1e-5), n


@pytest.mark.parametrize("name", CASES)
def test_oracle_greedy_generation_matches_reference(name):
    g = load(name)
    spec = g["spec"]
    cfg = O.OracleConfig(**spec["cfg"])
    sd = O.synthetic_state(cfg, seed=1)
    state = {}
    for step in range(spec["steps"]):  # generation in the fixture happens after the training steps
        x, y = batch_for(cfg, spec, step)
        _, _, grads = O.loss_and_grads(sd, cfg, x, y)
        c = O.clip_coef(O.grad_norm(grads), 1.0)
        O.adamw_step(sd, {k: v * c for k, v in grads.items()}, state, lr=spec["lr"], betas=tuple(spec["betas"]),
                     weight_decay=0.1, step=step + 1)
    prompt = torch.tensor(g["generate"]["prompt"])
    out = O.generate_greedy(sd, cfg, prompt, spec["gen_new"])
    assert out.tolist() == g["generate"]["tokens"]


@pytest.mark.parametrize("name", CASES)
def test_oracle_param_count_and_flops(name):
    g = load(name)
    cfg = O.OracleConfig(**g["spec"]["cfg"])
    assert O.num_params(cfg) == g["num_params"]
    # estimate_mfu(1, 1.0) = flops_per_token * T / 312e12
    assert O.flops_per_token(cfg) * cfg.block_size / 312e12 == pytest.approx(g["mfu_at_1s_per_iter"], rel=1e-9)


def test_survey_constants():
    # 6N + 12 L C T
    baby = O.OracleConfig(block_size=256, vocab_size=95, n_layer=6, n_head=6, n_embd=384)
    small = O.OracleConfig(block_size=1024, vocab_size=95, n_layer=12, n_head=12, n_embd=768)
    assert O.flops_per_token(baby) == 71_027_712
    assert O.flops_per_token(small) == 623_407_104
    assert O.num_params(small) == 85_026_816


def test_bf16_emulation_is_close_to_reference_autocast():
    g = load("tiny")
    spec = g["spec"]
    cfg = O.OracleConfig(**spec["cfg"])
    sd = O.synthetic_state(cfg, seed=1)
    x, y = batch_for(cfg, spec, 0)
    _, loss = O.forward(sd, cfg, x, y, bf16=True)
    assert abs(loss.item() - g["steps"][0]["bf16_autocast_loss"]) < 5e-3
    assert math.isfinite(loss.item())


def test_tunesformer_oracle_matches_reference_golden():
    """oracle/tunesformer_oracle.py against the loss / gradient norms the unmodified reference TunesFormer produced
    (oracle/make_golden_tunesformer.py)."""
    import json
    import os
    from oracle import tunesformer_oracle as TO
    from oracle.make_golden_tunesformer import inputs
    with open(os.path.join(os.path.dirname(__file__), "golden", "tunesformer_tiny.json")) as f:
        g = json.load(f)
    pc, cc, psd, csd, patches = inputs(g["spec"])
    loss, pg, cg = TO.loss_and_grads(psd, pc, csd, cc, patches)
    assert abs(loss.item() - g["reference"]["loss"]) <= 1e-6
    for grads, ref in ((pg, g["reference"]["patch_grad_norms"]), (cg, g["reference"]["char_grad_norms"])):
        for k, v in ref.items():
            assert abs(grads[k].norm().item() - v) <= 1e-4 * max(v, 1e-3) + 1e-7, (k, grads[k].norm().item(), v)
