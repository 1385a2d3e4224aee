"""One GroupNorm shape for ncu: python scripts/gn_one.py [C] [HW] [B]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from prompt_diffusion_b200 import ops
C = int(sys.argv[1]) if len(sys.argv) > 1 else 640
HW = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
B = int(sys.argv[3]) if len(sys.argv) > 3 else 16
x = torch.randn(B * HW, C, device="cuda").to(torch.bfloat16); o = torch.empty_like(x)
g, b = torch.randn(C, device="cuda"), torch.randn(C, device="cuda")
for _ in range(3): ops.group_norm(x, o, g, b, B, HW, eps=1e-5, act=1)
torch.cuda.synchronize(); print("ok")
