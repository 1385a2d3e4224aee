"""Per-unit phase timeline of CTA 0 of the persistent short-key tcgen05 attention kernel (engine 7; globaltimer stamps, us)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from prompt_diffusion_b200 import ops, _lib
B, h, N, Nk = 16, 8, int(sys.argv[2]) if len(sys.argv) > 2 else 4096, 77
d = int(sys.argv[1]) if len(sys.argv) > 1 else 40
C = h * d; dev = "cuda"
q = torch.randn(B * N, C, device=dev).to(torch.bfloat16)
kv = torch.randn(B * Nk, 2 * C, device=dev).to(torch.bfloat16)
out = torch.empty(B * N, C, device=dev, dtype=torch.bfloat16)
a = (q, kv[:, :C], kv[:, C:], out, B, h, N, Nk, d)
for _ in range(2): ops.attention(*a, engine=7)
dbg = torch.zeros(10 * 32, dtype=torch.int64, device=dev)
_lib.lib.pd_debug_attention_timeline(dbg.data_ptr())
ops.attention(*a, engine=7)
torch.cuda.synchronize(); _lib.lib.pd_debug_attention_timeline(None)
t = dbg.cpu().reshape(10, 32); t0 = int(t[t > 0].min())
print("unit | MMA: iter start  PV_a issued  PV_b issued | softmax A: S ready  P arrived | epilogue A: waits O  O ready  store issued  (us)")
for j in range(14):
    print("%4d | %10.2f %10.2f %10.2f | %10.2f %10.2f | %10.2f %10.2f %10.2f" % tuple([j] + [(int(t[k, j]) - t0) / 1e3 for k in range(8)]))
