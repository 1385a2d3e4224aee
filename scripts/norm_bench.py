"""CUDA-event micro-benchmark of the HBM-bound kernels (GroupNorm, LayerNorm, GEGLU) at the path's shapes.
Each shape runs over a ring of buffers larger than L2 ("cold") and on one buffer ("warm", L2-resident if it fits)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from prompt_diffusion_b200 import ops, _lib
dev = "cuda"
bf = torch.bfloat16
if os.environ.get("PD_GN_FUSED") == "0":      # A/B: statistics kernel + apply kernel (two PDL launches) instead of the cooperative kernel
    _lib.lib.pd_debug_group_norm_fused(0)

def timeit(fn, n, iters=20):
    """GPU time per call: the calls are captured into one CUDA graph (no host launch overhead in the number)."""
    for i in range(3): fn(i % n)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for i in range(iters): fn(i % n)
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    g.replay()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3

def ring(nbytes_one):
    return max(2, int(400e6 // nbytes_one) + 1)

print("kernel      B     HW     C |  cold us   GB/s | warm us   GB/s")
for (B, HW, C) in [(16, 4096, 320), (16, 4096, 640), (16, 4096, 960), (16, 1024, 640), (16, 1024, 1280), (16, 1024, 1920),
                   (16, 256, 1280), (16, 256, 2560), (16, 64, 1280), (16, 64, 2560)]:
    n = ring(B * HW * C * 2 * 2)
    xs = [torch.randn(B * HW, C, device=dev).to(bf) for _ in range(n)]
    os_ = [torch.empty_like(x) for x in xs]
    g, b = torch.randn(C, device=dev), torch.randn(C, device=dev)
    f = lambda i: ops.group_norm(xs[i], os_[i], g, b, B, HW, eps=1e-5, act=1)
    by = B * HW * C * 2 * 3       # read twice + write once (algorithmic, unfused)
    tc, tw = timeit(f, n), timeit(f, 1)
    print("gn      %5d %6d %5d | %8.1f %6.0f | %7.1f %6.0f" % (B, HW, C, tc, by / tc / 1e3, tw, by / tw / 1e3))
    del xs, os_
for (M, C) in [(65536, 320), (16384, 640), (4096, 1280), (1024, 1280)]:
    n = ring(M * C * 2 * 2)
    xs = [torch.randn(M, C, device=dev).to(bf) for _ in range(n)]
    os_ = [torch.empty_like(x) for x in xs]
    g, b = torch.randn(C, device=dev), torch.randn(C, device=dev)
    f = lambda i: ops.layer_norm(xs[i], os_[i], g, b)
    by = M * C * 2 * 2
    tc, tw = timeit(f, n), timeit(f, 1)
    print("ln      %5d %6d %5d | %8.1f %6.0f | %7.1f %6.0f" % (1, M, C, tc, by / tc / 1e3, tw, by / tw / 1e3))
    del xs, os_
for (M, F) in [(65536, 1280), (16384, 2560), (4096, 5120), (1024, 5120)]:
    n = ring(M * F * 2 * 3)
    xs = [torch.randn(M, 2 * F, device=dev).to(bf) for _ in range(n)]
    os_ = [torch.empty(M, F, device=dev, dtype=bf) for _ in range(n)]
    f = lambda i: ops.geglu(xs[i], os_[i])
    by = M * F * 2 * 3
    tc, tw = timeit(f, n), timeit(f, 1)
    print("geglu   %5d %6d %5d | %8.1f %6.0f | %7.1f %6.0f" % (1, M, F, tc, by / tc / 1e3, tw, by / tw / 1e3))
    del xs, os_
