"""One shape of pd_attention for ncu: python scripts/attn_one.py [d] [N] [B] [engine (default 3)]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from prompt_diffusion_b200 import ops
d = int(sys.argv[1]) if len(sys.argv) > 1 else 40
N = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
B = int(sys.argv[3]) if len(sys.argv) > 3 else 16
ENG = int(sys.argv[4]) if len(sys.argv) > 4 else 3
h = 8; C = h * d; dev = "cuda"
qkv = torch.randn(B * N, 3 * C, device=dev).to(torch.bfloat16)
out = torch.empty(B * N, C, device=dev, dtype=torch.bfloat16)
for _ in range(3): ops.attention(qkv[:, :C], qkv[:, C:2*C], qkv[:, 2*C:], out, B, h, N, N, d, engine=ENG)
torch.cuda.synchronize()
print("ok", float(out.float().abs().mean()))
