"""TunesFormer-shaped hierarchical decoder on the sm_100a kernels (BASELINE config 4).

Shape donor: tunesformer/utils.py:84-219 (paths relative to /root/reference) — a patch-level GPT-2 (bar patches of
PATCH_SIZE characters, embedded by Linear(PATCH_SIZE*128 -> n_embd) on one-hot rows, `GPT2Model(inputs_embeds=...)`) whose
hidden states become the FIRST input embedding of a character-level GPT-2 LM that spells out the next patch
(`GPT2LMHeadModel(inputs_embeds=cat(encoded_patch, wte(chars)[:, 1:]), labels=chars)`, pad id 0 ignored).

Both decoders are the GPT of model.py (biases, tanh GELU), so every matmul / attention / LayerNorm / loss / AdamW launch is
the same kernel as on the nanoGPT path; the joins are `GPT.forward_hidden` / `GPT.forward_with_first`.  The HF code itself
(third-party arithmetic, single-process DataParallel trainer) is not rebuilt; parameter names follow this package.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import ops
from .model import GPT, GPTConfig

PATCH_SIZE = 32      # tunesformer/config.py:1
CHAR_VOCAB = 128     # one-hot width per character (utils.py:102)


class _PatchEmbed(torch.autograd.Function):
    """out[m,:] = onehot(patch m) W^T + b as a GEMM on the one-hot rows; backward = weight / bias gradients only."""

    @staticmethod
    def forward(ctx, anchor, dec, patches):
        M = patches.shape[0]
        S, V, C = PATCH_SIZE, CHAR_VOCAB, dec.config.n_embd
        dec._ensure_device_state()
        oh = dec._onehot.get(M)
        if oh is None:
            oh = dec._onehot[M] = torch.empty(M, S * V, device=patches.device, dtype=torch.bfloat16)
        ops.onehot_bf16(patches.contiguous(), oh, V)
        out = torch.empty(M, C, device=patches.device, dtype=torch.float32)
        ops.gemm(oh, dec._view("shadow", "patch_embedding.weight"), epilogue=ops.EPI_F32, out=out,
                 bias=dec._view("flat", "patch_embedding.bias"))
        ctx.dec, ctx.oh = dec, oh
        return out

    @staticmethod
    def backward(ctx, d_out):
        dec, oh = ctx.dec, ctx.oh
        M, C = d_out.shape
        db16 = torch.empty(M, C, device=d_out.device, dtype=torch.bfloat16)
        ops.cast_bf16(d_out.contiguous().view(-1), db16.view(-1))
        # runs after the stack's backward plan (autograd order), which has zeroed / attached the gradient arena
        ops.gemm(db16, oh, a_mn=True, b_mn=True, epilogue=ops.EPI_F32_RED, out=dec._view("grad", "patch_embedding.weight"))
        ops.colsum_bf16(db16, dec._view("grad", "patch_embedding.bias"))
        sync = dec._grad_sync
        if sync is not None and dec.require_backward_grad_sync:
            sync.finalize()   # this arena's last gradients: wte/wpe, the patch embedding and the 1-D tail go out now
        return None, None, None


class PatchLevelDecoder(GPT):
    """GPT stack + patch embedding in ONE parameter arena (clip / AdamW / DDP buckets cover it without extra code)."""

    def __init__(self, config: GPTConfig):
        super().__init__(config)
        self.patch_embedding = nn.Linear(PATCH_SIZE * CHAR_VOCAB, config.n_embd)
        torch.nn.init.normal_(self.patch_embedding.weight, std=0.02)   # utils.py:94
        torch.nn.init.zeros_(self.patch_embedding.bias)
        self._onehot = {}
        self._flatten()

    def encode(self, patches):
        """patches int64 [B, P, PATCH_SIZE] -> fp32 [B, P, C] (last_hidden_state)."""
        B, P, S = patches.shape
        assert S == PATCH_SIZE and P <= self.config.block_size
        emb = _PatchEmbed.apply(self._anchor_tensor(patches.device), self, patches.reshape(B * P, S))
        return self.forward_hidden(emb.view(B, P, -1))


class TunesFormerShaped(nn.Module):
    def __init__(self, patch_config: GPTConfig, char_config: GPTConfig):
        super().__init__()
        assert char_config.vocab_size == CHAR_VOCAB and char_config.block_size >= PATCH_SIZE
        self.patch_level_decoder = PatchLevelDecoder(patch_config)
        self.char_level_decoder = GPT(char_config)
        self.pad_token_id = 0

    def forward(self, patches):
        """patches int64 [B, P, PATCH_SIZE] (pad id 0 at the tail of every patch) -> mean next-character loss over the
        non-pad characters of patches 1..P-1, each decoded from the encoding of the patches before it (utils.py:210-219)."""
        B, P, S = patches.shape
        encoded = self.patch_level_decoder.encode(patches)                  # [B, P, C]
        first = encoded[:, :-1, :].reshape(B * (P - 1), -1)
        chars = patches[:, 1:, :].reshape(B * (P - 1), S)
        y = torch.full_like(chars, -1)
        y[:, :-1] = chars[:, 1:]
        y[y == self.pad_token_id] = -1                                       # labels -100 at pads (utils.py:128-129)
        return self.char_level_decoder.forward_with_first(chars, first, y)

    def configure_optimizers(self, weight_decay, learning_rate, betas, device_type="cuda"):
        return _Optimizers([self.patch_level_decoder.configure_optimizers(weight_decay, learning_rate, betas, device_type),
                            self.char_level_decoder.configure_optimizers(weight_decay, learning_rate, betas, device_type)])

    def clip_grad_norm_(self, max_norm):
        """One global norm over both decoders, like torch.nn.utils.clip_grad_norm_(model.parameters(), max_norm)."""
        a, b = self.patch_level_decoder, self.char_level_decoder
        na, nb = a.clip_grad_norm_(max_norm), b.clip_grad_norm_(max_norm)
        total = na.square() + nb.square()
        for m in (a, b):   # both AdamW launches read the same device scalar: the squared global norm
            ss, mx = m._pending_clip
            ss.copy_(total.reshape(1))
        return total.sqrt()


class _Optimizers:
    """step / zero_grad / param_groups over the per-decoder FusedAdamW objects."""

    def __init__(self, opts):
        self.opts = opts

    @property
    def param_groups(self):
        return [g for o in self.opts for g in o.param_groups]

    def step(self):
        for o in self.opts:
            o.step()

    def zero_grad(self, set_to_none=True):
        for o in self.opts:
            o.zero_grad(set_to_none=set_to_none)

    def state_dict(self):
        return [o.state_dict() for o in self.opts]

    def load_state_dict(self, sds):
        for o, sd in zip(self.opts, sds):
            o.load_state_dict(sd)
