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

import random
import re

import numpy as np
import torch
import torch.nn as nn

from . import ops
from .model import GPT, GPTConfig

PATCH_SIZE = 32      # tunesformer/config.py:2
PATCH_LENGTH = 128   # tunesformer/config.py:1
CHAR_VOCAB = 128     # one-hot width per character (utils.py:102)


class Patchilizer:
    """Bar <-> patch codec of the hierarchical model (behaviour of tunesformer/utils.py:9-82; host-side, pure Python).

    A tune is cut at bar lines; every bar (with its closing delimiter) and every header line becomes one patch of PATCH_SIZE
    character codes: bos (1), the characters, eos (2), padded with 0.  (The reference passes the text through `unidecode`
    first; that package is optional here and plain ASCII input is unchanged by it.)"""

    DELIMS = ("|:", "::", ":|", "[|", "||", "|]", "|")
    pad_token_id, bos_token_id, eos_token_id = 0, 1, 2

    def __init__(self):
        self._split = re.compile("(" + "|".join(re.escape(d) for d in self.DELIMS) + ")")

    def split_bars(self, body):
        parts = [x for x in self._split.split("".join(body)) if x]
        if parts and parts[0] in self.DELIMS:       # a leading bar line belongs to the first bar
            parts[1] = parts[0] + parts[1]
            parts = parts[1:]
        return [parts[2 * i] + parts[2 * i + 1] for i in range(len(parts) // 2)]

    def bar2patch(self, bar, patch_size=PATCH_SIZE):
        codes = ([self.bos_token_id] + [ord(c) for c in bar] + [self.eos_token_id])[:patch_size]
        return codes + [self.pad_token_id] * (patch_size - len(codes))

    def patch2bar(self, patch):
        return "".join(chr(t) for t in patch if t > self.eos_token_id)

    def encode(self, abc_code, patch_length=PATCH_LENGTH, patch_size=PATCH_SIZE, add_special_patches=False):
        try:
            from unidecode import unidecode
            abc_code = unidecode(abc_code)
        except ImportError:
            pass
        patches, body = [], ""

        def flush(last_newline):
            bars = self.split_bars(body)
            for i, bar in enumerate(bars):
                patches.append(self.bar2patch(bar + "\n" if (last_newline and i == len(bars) - 1) else bar, patch_size))

        for line in (ln for ln in abc_code.split("\n") if ln):
            header = len(line) > 1 and ((line[0].isalpha() and line[1] == ":") or line.startswith("%%score"))
            if header:
                if body:
                    flush(True)
                    body = ""
                patches.append(self.bar2patch(line + "\n", patch_size))
            else:
                body += line + "\n"
        if body:
            flush(False)
        if add_special_patches:
            patches = ([[self.bos_token_id] * (patch_size - 1) + [self.eos_token_id]] + patches +
                       [[self.bos_token_id] + [self.eos_token_id] * (patch_size - 1)])
        return patches[:patch_length]

    def decode(self, patches):
        return "".join(self.patch2bar(p) for p in patches)


# ---- sampling helpers -------------------------------------------------------------------------------------------------
# The reference draws the next character with the third-party `samplings` package (top_p_sampling, top_k_sampling with
# return_probs=True, then temperature_sampling; tunesformer/utils.py:6,247-250).  The package is not vendored in the reference
# and is absent here, so these are restatements of its documented behaviour, not pinned to it: nucleus filter = keep the most
# probable characters up to and including the one that takes the cumulative probability to top_p, top-k filter = keep the k
# most probable (0 = all), both renormalise; the draw raises the probabilities to 1 / temperature, renormalises and samples
# with numpy's seeded generator.  With top_k = 1 every variant is the argmax, which is what the golden test pins.
def top_p_filter(probs, top_p):
    if top_p >= 1.0:
        return probs
    order = np.argsort(-probs, kind="stable")
    csum = np.cumsum(probs[order])
    keep = order[: int(np.searchsorted(csum, top_p, side="left")) + 1]
    out = np.zeros_like(probs)
    out[keep] = probs[keep]
    return out / out.sum()


def top_k_filter(probs, top_k):
    if top_k <= 0 or top_k >= probs.size:
        return probs
    keep = np.argsort(-probs, kind="stable")[:top_k]
    out = np.zeros_like(probs)
    out[keep] = probs[keep]
    return out / out.sum()


def temperature_draw(probs, temperature, seed=None):
    if temperature <= 0.0 or np.count_nonzero(probs) == 1:
        return int(np.argmax(probs))
    p = np.power(probs.astype(np.float64), 1.0 / temperature)
    p /= p.sum()
    return int(np.random.RandomState(seed).choice(p.size, p=p))


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

    @torch.no_grad()
    def generate(self, patches, tokens=None, top_p=1.0, top_k=0, temperature=1.0, seed=None):
        """One more patch for every tune of the batch (tunesformer/utils.py:221-255, there for a single tune).

        patches int64 [N, P, PATCH_SIZE] (or [P, PATCH_SIZE]): the tunes so far; `tokens` (optional, int64 [t] starting with
        bos): characters of the new patch that are already fixed (a prompt that ends inside a bar).  The patch-level decoder
        encodes the patches, the character-level decoder then spells the next patch one character at a time from the last
        patch's encoding — every step is one forward of all still-running tunes through the sm_100a kernels; a tune stops at
        eos or after PATCH_SIZE - 1 characters.  Returns (generated patches: one list of character codes per tune — a bare
        list when a single tune was given, like the reference —, the seed to pass to the next call)."""
        single = patches.dim() == 2 or patches.shape[0] == 1
        if patches.dim() == 2:
            patches = patches.unsqueeze(0)
        patches = patches.reshape(patches.shape[0], -1, PATCH_SIZE)
        N = patches.shape[0]
        dev = patches.device
        eos = 2
        first = self.patch_level_decoder.encode(patches)[:, -1, :].contiguous()           # [N, C]
        if tokens is None:
            tokens = torch.tensor([1], device=dev)
        seq = tokens.reshape(1, -1).to(dev).repeat(N, 1)                                    # [N, t]
        out = [[] for _ in range(N)]
        alive = list(range(N))
        rng = random.Random(seed)
        n_seed = None
        while alive:
            if seed is not None:                      # the reference re-seeds Python's generator with every draw
                n_seed = rng.randint(0, 1000000)
                rng.seed(n_seed)
            rows = torch.tensor(alive, device=dev)
            logits = self.char_level_decoder.next_logits_with_first(seq[rows].contiguous(), first[rows].contiguous())
            probs = torch.softmax(logits, dim=-1).cpu().numpy()
            nxt = []
            for j, i in enumerate(alive):
                p = top_k_filter(top_p_filter(probs[j], top_p), top_k)
                tok = temperature_draw(p, temperature, None if n_seed is None else n_seed + j)
                out[i].append(tok)
                nxt.append(tok)
            done = seq.shape[1] >= PATCH_SIZE - 1
            keep = [k for k, tok in enumerate(nxt) if tok != eos and not done]
            if not keep:
                break
            col = torch.zeros(N, 1, dtype=seq.dtype, device=dev)
            col[rows, 0] = torch.tensor(nxt, device=dev, dtype=seq.dtype)
            seq = torch.cat([seq, col], dim=1)
            alive = [alive[k] for k in keep]
        return (out[0] if single else out), n_seed

    @torch.no_grad()
    def generate_tune(self, prompt, max_patch=PATCH_LENGTH, top_p=0.8, top_k=8, temperature=1.2, seed=None):
        """The tune loop of tunesformer/generate.py:128-155 for one prompt (ABC header / opening bars as text): bar patches are
        generated and appended until the model emits an end patch, an empty bar, or `max_patch` patches.  Returns the text."""
        pz = Patchilizer()
        dev = next(self.parameters()).device
        patches = torch.tensor([pz.encode(prompt, add_special_patches=True)[:-1]], device=dev)
        prefix = pz.decode(patches[0].tolist())
        remaining = prompt[len(prefix):]
        tokens = torch.tensor([pz.bos_token_id] + [ord(c) for c in remaining], device=dev) if prompt else None
        tune = prompt
        while patches.shape[1] < max_patch:
            patch, seed = self.generate(patches, tokens, top_p=top_p, top_k=top_k, temperature=temperature, seed=seed)
            tokens = None
            if patch[0] == pz.eos_token_id:
                break
            bar = pz.decode([patch])
            if bar == "":
                break
            tune += bar
            nxt = torch.tensor([[pz.bar2patch(remaining + bar)]], device=dev)
            remaining = ""
            patches = torch.cat([patches, nxt], dim=1)
        return tune

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
