"""One cross-attention shape (Nk = 77) of pd_attention for ncu: python scripts/xattn_one.py [engine] [d] [N]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from prompt_diffusion_b200 import ops
eng = int(sys.argv[1]) if len(sys.argv) > 1 else 4
d = int(sys.argv[2]) if len(sys.argv) > 2 else 40
N = int(sys.argv[3]) if len(sys.argv) > 3 else 4096
B, h = 16, 8; C = h * d; dev = "cuda"
q = torch.randn(B * N, C, device=dev).to(torch.bfloat16)
kv = torch.randn(B * 77, 2 * C, device=dev).to(torch.bfloat16)
out = torch.empty(B * N, C, device=dev, dtype=torch.bfloat16)
for _ in range(3): ops.attention(q, kv[:, :C], kv[:, C:], out, B, h, N, 77, d, engine=eng)
torch.cuda.synchronize()
print("ok", float(out.float().abs().mean()))
