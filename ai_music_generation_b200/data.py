"""Device-resident token stream: the reference's `get_batch` (nanoGPT/train.py:122-144) without the per-batch host work.

The reference re-opens a np.memmap per batch, slices B windows in a Python list comprehension, widens them to int64, pins
and copies X and Y.  Here the whole `train.bin` / `val.bin` (flat uint16 ids; uint32 for the whitespace-tokenised corpus,
prepare_char.py:95-107) is uploaded to HBM once and one kernel gathers the B windows.  The window starts are drawn with
the same call the reference makes (`torch.randint(len(data) - block_size, (batch_size,))` on the CPU generator), so a run
seeded like the reference sees exactly the reference's batches.
"""
from __future__ import annotations

import os

import numpy as np
import torch

from . import ops


def check_vocab(tokens, vocab_size, path):
    """nn.Embedding raises on an id outside [0, V) (nanoGPT/model.py:177); the embedding kernel has no such check on its hot
    path, so a token file that does not match meta.pkl's vocab_size (or uint16 data read as uint32) is rejected here, once,
    when the file is loaded."""
    if vocab_size is None or len(tokens) == 0:
        return
    top = int(tokens.max())
    if top >= vocab_size:
        raise ValueError(f"{path}: token id {top} is outside the vocabulary (vocab_size = {vocab_size}); "
                         "wrong meta.pkl, or a token file of another width (uint16 / uint32)?")


class DeviceTokenStream:
    def __init__(self, data_dir, block_size, batch_size, device, wide_tokens=False, vocab_size=None):
        self.T, self.B, self.device = block_size, batch_size, torch.device(device)
        self.dtype = np.uint32 if wide_tokens else np.uint16
        self.data = {}
        for split in ("train", "val"):
            path = os.path.join(data_dir, f"{split}.bin")
            if os.path.exists(path):
                host = np.fromfile(path, dtype=self.dtype)
                check_vocab(host, vocab_size, path)
                # torch has no uint32 tensors for arithmetic; the kernel only needs the bytes
                view = host.view(np.int16 if self.dtype == np.uint16 else np.int32)
                self.data[split] = torch.from_numpy(view).to(self.device)

    def __len__(self):
        return self.data["train"].numel()

    def get(self, split):
        data = self.data["train" if split == "train" else "val"]
        ix = torch.randint(data.numel() - self.T, (self.B,))  # same draw as the reference
        ix_dev = ix.pin_memory().to(self.device, non_blocking=True)  # fresh pinned staging per call: the copy is async
        x = torch.empty(self.B, self.T, device=self.device, dtype=torch.int64)
        y = torch.empty(self.B, self.T, device=self.device, dtype=torch.int64)
        ops.sample_batch(data, ix_dev, x, y)
        return x, y
