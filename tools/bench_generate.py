"""cfg5 (BASELINE.json configs[4]): batched generation, GPT-2-small shape, 256 tunes x 1024 new ABC tokens, greedy.
Prompt = the single start token (sample.py:32), so the window never slides and the KV-cache path covers every token."""
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ai_music_generation_b200 import GPT, GPTConfig, ops  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
N = int(sys.argv[2]) if len(sys.argv) > 2 else 1023
torch.manual_seed(1337)
model = GPT(GPTConfig(block_size=1024, vocab_size=95, n_layer=12, n_head=12, n_embd=768, dropout=0.0, bias=False)).cuda().eval()
x = torch.zeros(B, 1, dtype=torch.long, device="cuda")
model.generate(x, 8, top_k=1)
torch.cuda.synchronize()
l0 = ops.LAUNCHES
t0 = time.time()
y = model.generate(x, N, top_k=1)
torch.cuda.synchronize()
dt = time.time() - t0
print(json.dumps({"workload": f"generate {B} x {N} new tokens, greedy, KV cache", "seconds": dt, "tokens_per_s": B * N / dt,
                  "ms_per_token_step": dt / N * 1e3, "launches": ops.LAUNCHES - l0}))
t0 = time.time()
y2 = model.generate(x[:8], 64, top_k=1, use_cache=False)
torch.cuda.synchronize()
print(json.dumps({"workload": "generate 8 x 64 new tokens, reference-style context recompute", "seconds": time.time() - t0}))
