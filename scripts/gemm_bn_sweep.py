"""Tile-shape sweep for the small-M (8x8 latent) layers: time vs forced BN / CTA group (split-K planning)."""
import math, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from prompt_diffusion_b200 import ops, _lib
from prompt_diffusion_b200._lib import PD_ENGINE_TC
dev = "cuda"
def run(B, H, W, C, N, ks, iters=20):
    M = B * H * W
    x = torch.randn(M, C, device=dev).to(torch.bfloat16)
    w = (torch.randn(N, ks * ks * C, device=dev) / math.sqrt(ks * ks * C)).to(torch.bfloat16)
    bias = torch.randn(N, device=dev); out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    f = lambda: ops.conv2d(x, w, out, B, H, W, ksize=ks, bias=bias, engine=PD_ENGINE_TC)
    for _ in range(3): f()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(iters): f()
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3
print("shape                      cg  BN |   us")
for (B, H, W, C, N, ks) in [(16, 8, 8, 1280, 1280, 3), (16, 8, 8, 1280 // 3 // 64 * 64, 1280, 3), (16, 16, 16, 1280, 1280, 3)]:
    for cg in (1, 2):
        for bn in (96, 128, 160, 256):
            _lib.lib.pd_debug_force_cta_group(cg); _lib.lib.pd_debug_force_bn(bn)
            us = run(B, H, W, C, N, ks)
            print("B%d %dx%d C%d N%d k%d   %d %4d | %7.1f" % (B, H, W, C, N, ks, cg, bn, us))
_lib.lib.pd_debug_force_cta_group(0); _lib.lib.pd_debug_force_bn(0)
