"""Per-tile phase timeline of CTA (0,0,0) of the four-group tcgen05 attention kernel (globaltimer stamps, us)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from prompt_diffusion_b200 import ops, _lib
B, h, N, d = 16, 8, 4096, 40
ENG = int(sys.argv[1]) if len(sys.argv) > 1 else 5
C = h * d; dev = "cuda"
qkv = torch.randn(B * N, 3 * C, device=dev).to(torch.bfloat16)
out = torch.empty(B * N, C, device=dev, dtype=torch.bfloat16)
for _ in range(2): ops.attention(qkv[:, :C], qkv[:, C:2*C], qkv[:, 2*C:], out, B, h, N, N, d, engine=ENG)
dbg = torch.zeros(12 * 32, dtype=torch.int64, device=dev)
_lib.lib.pd_debug_attention_timeline(dbg.data_ptr())
ops.attention(qkv[:, :C], qkv[:, C:2*C], qkv[:, 2*C:], out, B, h, N, N, d, engine=ENG)
torch.cuda.synchronize(); _lib.lib.pd_debug_attention_timeline(None)
t = dbg.cpu().reshape(12, 32); t0 = int(t[t > 0].min())
print("tile | MMA saw p_full g0 g1 g2 g3 | S in regs g0 g1 g2 g3 | arrived g0 g1 g2 g3   (us)")
for j in range(16):
    print("%4d | " % j + " ".join("%7.2f" % ((int(t[k, j]) - t0) / 1e3) for k in range(4)) + " | " +
          " ".join("%7.2f" % ((int(t[k, j]) - t0) / 1e3) for k in range(4, 8)) + " | " +
          " ".join("%7.2f" % ((int(t[k, j]) - t0) / 1e3) for k in range(8, 12)))
