# GPT-2-small shape (12L/12H/768d, block 1024) on character-level ABC, B200-sized micro-batch (BASELINE.json configs[2]);
# optimiser recipe of the reference's nanoGPT/config/train_gpt2.py:11-25
out_dir = "out-abc-gpt2-small"
dataset = "irishman"
batch_size = 32
block_size = 1024
gradient_accumulation_steps = 8
n_layer = 12
n_head = 12
n_embd = 768
dropout = 0.0
max_iters = 600000
lr_decay_iters = 600000
eval_interval = 1000
eval_iters = 200
log_interval = 10
weight_decay = 1e-1
