"""Pins oracle/tunesformer_oracle.py to the UNMODIFIED reference TunesFormer (tunesformer/utils.py, HF GPT-2 classes).

Run in the build container (needs /root/reference and `transformers`):  python oracle/make_golden_tunesformer.py
Writes tests/golden/tunesformer_tiny.json: the reference's loss and per-tensor gradient norms for closed-form weights and a
seeded patch tensor (fixtures hold outputs only).  `unidecode` and `samplings` (text clean-up / sampling helpers that the
forward pass never touches) are absent here and are stubbed so that the module imports.
"""
import json
import os
import sys
import types

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import nanogpt_oracle as O  # noqa: E402

SPEC = dict(
    patch_cfg=dict(block_size=16, vocab_size=1, n_layer=2, n_head=2, n_embd=128, dropout=0.0, bias=True, activation="gelu_tanh"),
    char_cfg=dict(block_size=32, vocab_size=128, n_layer=2, n_head=2, n_embd=128, dropout=0.0, bias=True, activation="gelu_tanh"),
    patch_seed=5, char_seed=6, data_seed=0, n_patches=12,
)


GEN_WTE_SCALE = 40.0


def inputs(spec):
    pc, cc = O.OracleConfig(**spec["patch_cfg"]), O.OracleConfig(**spec["char_cfg"])
    psd, csd = O.synthetic_state(pc, seed=spec["patch_seed"]), O.synthetic_state(cc, seed=spec["char_seed"])
    g = torch.Generator().manual_seed(spec["data_seed"])
    psd["patch_embedding.weight"] = torch.randn(pc.n_embd, 32 * 128, generator=g) * 0.02
    psd["patch_embedding.bias"] = torch.randn(pc.n_embd, generator=g) * 0.02
    patches = torch.randint(3, 128, (1, spec["n_patches"], 32), generator=g)
    lens = torch.randint(6, 33, (1, spec["n_patches"]), generator=g)
    patches[torch.arange(32)[None, None, :] >= lens[..., None]] = 0
    return pc, cc, psd, csd, patches


def to_hf(sd, prefix):
    """our names -> HF GPT-2 names; Conv1D stores [in, out]"""
    out = {}
    for k, v in sd.items():
        if k.startswith("patch_embedding"):
            continue
        hk = k.replace("transformer.", "")
        if any(hk.endswith(s) for s in ("c_attn.weight", "c_proj.weight", "c_fc.weight")):
            v = v.t().contiguous()
        out[prefix + hk] = v.clone()
    return out


def main():
    for name in ("unidecode", "samplings"):
        m = types.ModuleType(name)
        m.unidecode = lambda s: s
        m.top_p_sampling = m.top_k_sampling = m.temperature_sampling = None
        sys.modules[name] = m
    sys.path.insert(0, "/root/reference/tunesformer")
    import utils as ref  # the reference module, unmodified
    from transformers import GPT2Config

    pc, cc, psd, csd, patches = inputs(SPEC)
    kw = dict(n_embd=128, n_head=2, resid_pdrop=0.0, embd_pdrop=0.0, attn_pdrop=0.0)
    patch_config = GPT2Config(num_hidden_layers=pc.n_layer, max_length=pc.block_size, max_position_embeddings=pc.block_size,
                              vocab_size=1, **kw)
    char_config = GPT2Config(num_hidden_layers=cc.n_layer, max_length=cc.block_size, max_position_embeddings=cc.block_size,
                             vocab_size=128, **kw)
    model = ref.TunesFormer(patch_config, char_config, share_weights=False)
    model.train()
    p_hf = to_hf(psd, "patch_level_decoder.base.")
    p_hf["patch_level_decoder.patch_embedding.weight"] = psd["patch_embedding.weight"].clone()
    p_hf["patch_level_decoder.patch_embedding.bias"] = psd["patch_embedding.bias"].clone()
    c_hf = to_hf(csd, "char_level_decoder.base.transformer.")
    c_hf["char_level_decoder.base.lm_head.weight"] = csd["transformer.wte.weight"].clone()
    missing, unexpected = model.load_state_dict({**p_hf, **c_hf}, strict=False)
    assert not unexpected, unexpected
    assert all("attn.bias" in k or "masked_bias" in k for k in missing), missing
    out = model(patches.reshape(1, -1), 0)
    loss = out.loss
    loss.backward()
    named = dict(model.named_parameters())
    rec = {"loss": loss.item(), "patch_grad_norms": {}, "char_grad_norms": {}}
    for k in psd:
        hk = "patch_level_decoder." + (k if k.startswith("patch_embedding") else "base." + k.replace("transformer.", ""))
        gr = named[hk].grad
        rec["patch_grad_norms"][k] = 0.0 if gr is None else gr.norm().item()
    for k in csd:
        hk = "char_level_decoder.base.transformer." + k.replace("transformer.", "")
        rec["char_grad_norms"][k] = named[hk].grad.norm().item()
    path = os.path.join(ROOT, "tests", "golden", "tunesformer_tiny.json")
    with open(path, "w") as f:
        json.dump({"spec": SPEC, "reference": rec, "transformers": __import__("transformers").__version__}, f, indent=1)
    print("wrote", path, "loss", rec["loss"])

    # ---- generation (tunesformer/utils.py:221-255) under a GREEDY sampler: `samplings` is absent, so the three helpers the
    # reference imports from it are stubbed as pass-through filters + argmax; the fixture pins everything else of the loop
    # (patch encoding, first-embedding substitution, stop rules) and the character-level logits through their arg-max / margins.
    import numpy as np
    ref.top_p_sampling = lambda prob, top_p=1, return_probs=True: prob
    ref.top_k_sampling = lambda prob, top_k=0, return_probs=True: prob
    margins = []

    def greedy(prob, temperature=1, seed=None):
        o = np.sort(prob)[::-1]
        margins.append(float(o[0] - o[1]))
        return int(np.argmax(prob))

    ref.temperature_sampling = greedy
    model.eval()
    # closed-form weights give nearly flat next-character distributions (top-2 margins ~1e-5): for a decision that survives bf16
    # arithmetic the tied character embedding / head matrix is scaled up (GEN_WTE_SCALE), which sharpens the logits
    with torch.no_grad():
        model.char_level_decoder.base.transformer.wte.weight.mul_(GEN_WTE_SCALE)
        model.char_level_decoder.base.lm_head.weight.copy_(model.char_level_decoder.base.transformer.wte.weight)
    pz = ref.Patchilizer()
    seq = patches[:, :5, :].clone()
    gen = []
    with torch.no_grad():
        for _ in range(4):
            patch, _ = model.generate(seq.reshape(1, -1), None, top_p=1, top_k=0, temperature=1, seed=None)
            gen.append([int(t) for t in patch])
            if patch[0] == pz.eos_token_id:
                break
            bar = pz.decode([patch])
            if bar == "":
                break
            seq = torch.cat([seq, torch.tensor([[pz.bar2patch(bar)]])], dim=1)
        fixed = torch.tensor([1, 70, 71])     # a prompt that ends inside a bar: bos + two fixed characters
        patch2, _ = model.generate(patches[:, :5, :].reshape(1, -1), fixed, top_p=1, top_k=0, temperature=1, seed=None)
    texts = ["X:1\nL:1/8\nM:3/4\nK:D\n de |\"D\" fa fd AF | \"G\" GB dB GB |]\n",
             "S:2\nB:9\nE:4\nB:9\nL:1/8\nM:3/4\nK:D\n de |\"D\" ",
             "|: A2 B2 :: c4 d4 e4 f4 g4 a4 b4 c'4 d'4 e'4 f'4 :|\n%%score 1 2\nV:1\n[| z8 || x | y |]",
             "K:G\nabc"]
    codec = [{"text": t, "plain": pz.encode(t), "special": pz.encode(t, add_special_patches=True),
              "decoded": pz.decode(pz.encode(t, add_special_patches=True))} for t in texts]
    path = os.path.join(ROOT, "tests", "golden", "tunesformer_tiny_generate.json")
    with open(path, "w") as f:
        json.dump({"spec": SPEC, "n_prompt_patches": 5, "generated": gen, "margins": margins, "with_fixed_tokens": [int(t) for t in patch2],
                   "fixed_tokens": fixed.tolist(), "char_wte_scale": GEN_WTE_SCALE, "codec": codec, "sampler": "greedy stub (samplings absent)"}, f, indent=1)
    print("wrote", path, "patches", [len(g) for g in gen], "min margin", min(margins))


if __name__ == "__main__":
    main()
