import os, sys, torch
sys.path.insert(0, "/root/repo")
from ai_music_generation_b200 import ops
B,T,H=32,1024,12; C=H*64
torch.manual_seed(0)
qkv=torch.randn(B*T,3*C,device="cuda").bfloat16(); out=torch.zeros(B*T,C,device="cuda",dtype=torch.bfloat16); lse=torch.zeros(B,H,T,device="cuda")
dout=torch.randn(B*T,C,device="cuda").bfloat16(); dqkv=torch.zeros(B*T,3*C,device="cuda",dtype=torch.bfloat16); delta=torch.zeros(B,H,T,device="cuda")
for _ in range(int(sys.argv[1]) if len(sys.argv)>1 else 3):
    ops.attn_fwd(qkv,out,lse,B,T,H); ops.attn_bwd(qkv,out,dout,lse,delta,dqkv,B,T,H)
torch.cuda.synchronize()
