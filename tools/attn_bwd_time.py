import os, sys, torch
sys.path.insert(0, os.getcwd())
from ai_music_generation_b200 import ops
B,T,H=32,1024,12; C=H*64
qkv=torch.randn(B*T,3*C,device="cuda").bfloat16(); o=torch.empty(B*T,C,device="cuda",dtype=torch.bfloat16)
lse=torch.empty(B,H,T,device="cuda"); do=torch.randn(B*T,C,device="cuda").bfloat16()
dqkv=torch.empty(B*T,3*C,device="cuda",dtype=torch.bfloat16); delta=torch.empty(B,H,T,device="cuda")
ops.attn_fwd(qkv,o,lse,B,T,H)
for _ in range(3): ops.attn_bwd(qkv,o,do,lse,delta,dqkv,B,T,H)
torch.cuda.synchronize()
e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20): ops.attn_bwd(qkv,o,do,lse,delta,dqkv,B,T,H)
e1.record(); torch.cuda.synchronize(); print("bwd ms", e0.elapsed_time(e1)/20)
