"""Dropout (all music configs of the reference train with dropout 0.2, config/irishman_char.py:25): the fused masks are
counter-based, so the exact masks the kernels applied are regenerated on the host (ai_music_generation_b200/dropout.py)
and handed to the CPU oracle; forward and backward must then agree like in the dropout-free parity tests."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import nanogpt_oracle as O

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def host_masks(cfg, B, T, p, seed):
    from ai_music_generation_b200 import dropout as D
    L, C, H = cfg.n_layer, cfg.n_embd, cfg.n_head
    t = lambda a: torch.from_numpy(a)  # noqa: E731
    m = {"p": p, "emb": t(D.keep_mask(D.site_key(seed, 0), B * T, C, p)).view(B, T, C), "attn_p": [], "attn_resid": [], "mlp_resid": []}
    for l in range(L):
        m["attn_p"].append(t(D.attention_keep_mask(D.site_key(seed, 1 + 3 * l), B, H, T, p)))
        m["attn_resid"].append(t(D.keep_mask(D.site_key(seed, 2 + 3 * l), B * T, C, p)).view(B, T, C))
        m["mlp_resid"].append(t(D.keep_mask(D.site_key(seed, 3 + 3 * l), B * T, C, p)).view(B, T, C))
    return m


@pytest.mark.parametrize("name,p", [("tiny", 0.2), ("tiny_bias", 0.1), ("baby", 0.2)])
def test_dropout_forward_backward_match_oracle_with_same_masks(name, p, cuda_device):
    from ai_music_generation_b200 import GPT, GPTConfig
    with open(os.path.join(GOLDEN, f"nanogpt_{name}.json")) as f:
        spec = json.load(f)["spec"]
    _check_against_oracle(dict(spec["cfg"], dropout=p), spec["batch"], spec["seqlen"], p, cuda_device)


def test_dropout_on_packed_short_sequences(cuda_device):
    """T = 32 with a batch that fills whole 128-row tiles: the attention kernels pack four sequences per tile (block-
    diagonal masking); masks and row statistics keep the canonical per-sequence counters, so the host twin is unchanged."""
    cfgd = dict(block_size=32, vocab_size=95, n_layer=2, n_head=2, n_embd=128, dropout=0.2, bias=True)
    _check_against_oracle(cfgd, 8, 32, 0.2, cuda_device)


def _check_against_oracle(cfgd, B, T, p, cuda_device):
    from ai_music_generation_b200 import GPT, GPTConfig
    cfg = O.OracleConfig(**cfgd)
    sd = O.synthetic_state(cfg, seed=1)
    model = GPT(GPTConfig(**cfgd))
    model.load_state_dict({**sd, "lm_head.weight": sd["transformer.wte.weight"]})
    model = model.to(cuda_device).train()
    x, y = O.synthetic_tokens(cfg, B, T, seed=0)
    model._next_dropout_seed = 4242
    logits, loss = model(x.to(cuda_device), y.to(cuda_device))
    assert model.last_dropout_seed == 4242
    loss.backward()
    masks = host_masks(cfg, B, T, p, 4242)
    torch.set_num_threads(8)
    ref_loss, ref_logits, ref_grads = O.loss_and_grads(sd, cfg, x, y, bf16=True, masks=masks)
    tru_loss, tru_logits, tru_grads = O.loss_and_grads(sd, cfg, x, y, bf16=False, masks=masks)
    assert abs(loss.item() - ref_loss.item()) <= 3e-3, (loss.item(), ref_loss.item(), tru_loss.item())
    ours = (logits.float().cpu() - tru_logits).abs()
    refs = (ref_logits - tru_logits).abs()
    assert ours.max().item() <= 1.5 * refs.max().item() + 1e-2
    assert ours.mean().item() <= 1.5 * refs.mean().item() + 1e-3
    named = dict(model.named_parameters())
    for n, tg in tru_grads.items():
        got = named[n].grad.float().cpu()
        ours_rel = ((got - tg).norm() / tg.norm()).item()
        ref_rel = ((ref_grads[n] - tg).norm() / tg.norm()).item()
        assert ours_rel <= 1.5 * ref_rel + 5e-3, (n, ours_rel, ref_rel)
    # and the dropout really happened: the dropout-free loss is different
    nodrop_loss, _, _ = O.loss_and_grads(sd, cfg, x, y, bf16=True)
    assert abs(nodrop_loss.item() - loss.item()) > 1e-3


def test_dropout_seed_semantics_and_eval_mode(cuda_device):
    from ai_music_generation_b200 import GPT, GPTConfig
    cfgd = dict(block_size=64, vocab_size=95, n_layer=2, n_head=2, n_embd=128, dropout=0.2, bias=False)
    torch.manual_seed(0)
    model = GPT(GPTConfig(**cfgd)).to(cuda_device)
    x = torch.randint(95, (4, 64), device=cuda_device)
    y = torch.randint(95, (4, 64), device=cuda_device)
    model.train()
    torch.manual_seed(7)
    _, l1 = model(x, y)
    a = l1.item()
    torch.manual_seed(7)
    _, l2 = model(x, y)
    assert l2.item() == a                       # same torch seed -> same masks
    _, l3 = model(x, y)
    assert l3.item() != a                       # the seed advances from call to call
    model.eval()
    with torch.no_grad():
        _, e1 = model(x, y)
        _, e2 = model(x, y)
    assert e1.item() == e2.item() and model.last_dropout_seed is None


def test_dropout_keep_rate(cuda_device):
    from ai_music_generation_b200 import dropout as D, ops
    idx = torch.zeros(64 * 128, dtype=torch.int64, device=cuda_device)
    wte = torch.ones(4, 256, device=cuda_device)
    wpe = torch.zeros(128, 256, device=cuda_device)
    out = torch.empty(64 * 128, 256, device=cuda_device)
    key = D.site_key(99, 0)
    ops.embed_fwd(idx, wte, wpe, out, 128, drop_p=0.2, drop_key=key)
    kept = (out != 0)
    assert abs(kept.float().mean().item() - 0.8) < 5e-3
    assert torch.allclose(out[kept], torch.full_like(out[kept], 1.25))
    host = torch.from_numpy(D.keep_mask(key, 64 * 128, 256, 0.2)).to(cuda_device)
    assert torch.equal(kept, host)              # bit-exact agreement between the device hash and its host twin
