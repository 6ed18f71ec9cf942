"""CPU oracle for the TunesFormer-shaped hierarchical decoder — TEST INFRASTRUCTURE ONLY (see nanogpt_oracle.py).

Restates tunesformer/utils.py:84-219 (paths relative to /root/reference) on top of the nanoGPT oracle's functional
forward: PatchLevelDecoder.forward (:96-106: one-hot rows -> Linear(32*128 -> C) -> GPT-2 stack with inputs_embeds ->
last_hidden_state), CharLevelDecoder.forward (:120-154: pad id 0 -> label -100, first input embedding = encoded patch, HF
shifted cross-entropy) and TunesFormer.forward (:210-219: patch i decodes from the encoding of patches <= i-1... i.e.
encoded[:-1] against patches[1:]).  GPT-2 blocks = pre-LN blocks with biases and the tanh GELU, which is what
OracleConfig(bias=True, activation="gelu_tanh") runs.

Pinned: oracle/make_golden_tunesformer.py imports the UNMODIFIED reference module (with the installed `transformers`
GPT-2 classes it is built on; Conv1D [in, out] weights transposed from this layout), runs TunesFormer.forward + backward on
closed-form weights and commits its loss and per-tensor gradient norms as tests/golden/tunesformer_tiny.json;
tests/test_oracle_golden.py holds this restatement to them (loss 1e-6, gradient norms 1e-4).
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

from . import nanogpt_oracle as O

PATCH_SIZE, CHAR_VOCAB = 32, 128


def forward(patch_sd, patch_cfg, char_sd, char_cfg, patches, bf16=False):
    """patches int64 [B, P, 32] -> scalar loss."""
    B, P, S = patches.shape
    oh = F.one_hot(patches, CHAR_VOCAB).float().reshape(B, P, S * CHAR_VOCAB)
    w, b = patch_sd["patch_embedding.weight"], patch_sd["patch_embedding.bias"]
    emb = (O._bf(oh) @ O._bf(w).t() + b) if bf16 else (oh @ w.t() + b)       # fp32 output of a bf16 GEMM under autocast
    _, _, hidden = O.forward(patch_sd, patch_cfg, None, None, bf16=bf16, return_hidden=True, inputs_embeds=emb)
    first = hidden[:, :-1, :].reshape(B * (P - 1), -1)
    chars = patches[:, 1:, :].reshape(B * (P - 1), S)
    y = torch.full_like(chars, -1)
    y[:, :-1] = chars[:, 1:]
    y[y == 0] = -1
    _, loss = O.forward(char_sd, char_cfg, chars, y, bf16=bf16, first_embeds=first)
    return loss


def loss_and_grads(patch_sd, patch_cfg, char_sd, char_cfg, patches, bf16=False):
    pl = {k: v.detach().clone().requires_grad_(True) for k, v in patch_sd.items()}
    cl = {k: v.detach().clone().requires_grad_(True) for k, v in char_sd.items()}
    loss = forward(pl, patch_cfg, cl, char_cfg, patches, bf16=bf16)
    loss.backward()
    z = lambda d: {k: (v.grad if v.grad is not None else torch.zeros_like(v)) for k, v in d.items()}  # noqa: E731
    return loss.detach(), z(pl), z(cl)
