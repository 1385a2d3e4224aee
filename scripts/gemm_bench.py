"""Micro-benchmark of pd_conv2d (tcgen05 engine) on single layer shapes, CUDA-event timed.
    python scripts/gemm_bench.py                 # table over the path's dominant shapes
    python scripts/gemm_bench.py --one M N K ks res   # one shape, few launches (for ncu)
"""
import argparse, math, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from prompt_diffusion_b200 import ops
from prompt_diffusion_b200._lib import PD_ENGINE_TC

dev = "cuda"
BLOCKED = False
def run(B, H, W, C, N, ks, res, iters=20, warm=3, rowvec=False):
    M = B * H * W
    x = torch.randn(M, C, device=dev).to(torch.bfloat16)
    w = (torch.randn(N, ks * ks * C, device=dev) / math.sqrt(ks * ks * C)).to(torch.bfloat16)
    if BLOCKED:
        w = ops.block_weight(w)
    bias = torch.randn(N, device=dev)
    out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    r = torch.randn(M, N, device=dev).to(torch.bfloat16) if res else None
    rv = torch.randn(B, N, device=dev) if rowvec else None
    # rotate over enough distinct buffers that nothing stays L2-resident between launches
    for _ in range(warm):
        ops.conv2d(x, w, out, B, H, W, ksize=ks, bias=bias, res=r, rowvec=rv, engine=PD_ENGINE_TC)
    torch.cuda.synchronize()
    # GPU time per launch: the launches are captured into one CUDA graph (no host launch overhead in the number)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(iters):
            ops.conv2d(x, w, out, B, H, W, ksize=ks, bias=bias, res=r, rowvec=rv, engine=PD_ENGINE_TC)
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    g.replay()
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / iters * 1e3
    fl = 2.0 * M * N * ks * ks * C
    by = (M * C + M * N * (2 if res else 1) + N * ks * ks * C) * 2
    return us, fl / us / 1e6, by / us / 1e3

ap = argparse.ArgumentParser()
ap.add_argument("--one", nargs=7, type=int, default=None, help="B H W C N ks res")
ap.add_argument("--iters", type=int, default=20)
ap.add_argument("--blocked", action="store_true", help="k-block-major weight layout")
ap.add_argument("--small", action="store_true", help="fixed-overhead study: tiny and short-K GEMMs")
ap.add_argument("--modes", action="store_true", help="timing experiment: full kernel vs no-MMA vs no-TMA, per tile shape")
ap.add_argument("--epi-modes", action="store_true", help="timing experiment on the epilogue of short-K layers (gemm_sm100.cu dbg_mode 3..7)")
ap.add_argument("--bres", action="store_true", help="resident-B schedule against the tabled variant on the short-K layer shapes")
ap.add_argument("--sk", action="store_true", help="stream-K tile shapes on the small-M (8x8 / 16x16 level) 3x3 layers")
ap.add_argument("--vh", action="store_true", help="vertical-halo schedule against the tabled variant on the 3x3 layer shapes")
a = ap.parse_args()
BLOCKED = a.blocked
if a.vh:
    from prompt_diffusion_b200 import _lib
    L = _lib.lib
    BLOCKED = True
    shapes = [(16, 64, 64, 320, 320, 3, 0), (16, 64, 64, 640, 320, 3, 0), (16, 64, 64, 960, 320, 3, 0), (16, 32, 32, 640, 640, 3, 0),
              (16, 32, 32, 1280, 640, 3, 0), (16, 32, 32, 320, 640, 3, 0), (16, 16, 16, 1280, 1280, 3, 0), (16, 16, 16, 2560, 1280, 3, 0)]
    combos = [(1, 0), (2, 0), (2, 160), (2, 128), (2, 256)]
    print("   B   HxW     C     N ks res | tabled us  TFLOP/s | " + " | ".join("vh cg%d bn%-3d" % c for c in combos))
    for s_ in shapes:
        base, tf, _ = run(*s_, iters=a.iters)
        row = []
        for cg, bn in combos:
            if bn > s_[4]: row.append(None); continue
            L.pd_debug_force_cta_group(cg); L.pd_debug_force_bn(bn); L.pd_debug_force_vh(1)
            n0 = L.pd_debug_vh_launches()
            try:
                us = run(*s_, iters=a.iters)[0]
            except RuntimeError:
                us = None
            took = L.pd_debug_vh_launches() > n0
            L.pd_debug_force_cta_group(0); L.pd_debug_force_bn(0); L.pd_debug_force_vh(0)
            row.append(us if took else None)
        print("%4d %3dx%-3d %5d %5d %2d %3d | %9.1f %8.0f | " % (*s_, base, tf) + " | ".join("%12.1f" % r if r is not None else "           -" for r in row))
elif a.sk:
    from prompt_diffusion_b200 import _lib
    L = _lib.lib
    BLOCKED = True
    shapes = [(16, 8, 8, 1280, 1280, 3, 1), (16, 8, 8, 2560, 1280, 3, 0), (16, 8, 8, 1280, 1280, 1, 1), (16, 16, 16, 1280, 1280, 3, 1), (16, 16, 16, 2560, 1280, 3, 0)]
    combos = [(1, 128, 1), (1, 256, 1), (2, 128, 1), (2, 256, 1), (1, 128, 0), (2, 256, 0)]
    print("   B   HxW     C     N ks res | tabled us | " + " | ".join("cg%d bn%-3d sk%d" % c for c in combos))
    for s_ in shapes:
        base = run(*s_, iters=a.iters)[0]
        row = []
        for cg, bn, sk in combos:
            L.pd_debug_force_cta_group(cg); L.pd_debug_force_bn(bn); L.pd_debug_force_stream_k(sk)
            try:
                us = run(*s_, iters=a.iters)[0]
            except RuntimeError:
                us = float("nan")
            L.pd_debug_force_cta_group(0); L.pd_debug_force_bn(0); L.pd_debug_force_stream_k(0)
            row.append(us)
        print("%4d %3dx%-3d %5d %5d %2d %3d | %9.1f | " % (*s_, base) + " | ".join("%13.1f" % r for r in row))
elif a.bres:
    from prompt_diffusion_b200 import _lib
    L = _lib.lib
    BLOCKED = True     # the packed model weights are k-block-major
    shapes = [(16, 64, 64, 320, 320, 1, 1), (16, 64, 64, 320, 320, 1, 0), (16, 64, 64, 320, 960, 1, 0), (16, 64, 64, 320, 2560, 1, 0),
              (16, 32, 32, 640, 640, 1, 1), (16, 32, 32, 640, 640, 1, 0), (16, 32, 32, 640, 1920, 1, 0), (16, 32, 32, 640, 5120, 1, 0),
              (16, 16, 16, 1280, 1280, 1, 1)]
    combos = [(1, 96), (1, 128), (1, 160), (2, 128), (2, 160), (2, 192), (2, 256)]
    print("   B   HxW     C     N ks res | tabled us | " + " | ".join("cg%d bn%-3d" % c for c in combos) + "   (resident B; '-' = does not fit / apply)")
    for s_ in shapes:
        base = run(*s_, iters=a.iters)[0]
        row = []
        for cg, bn in combos:
            if s_[4] < bn: row.append(None); continue
            L.pd_debug_force_cta_group(cg); L.pd_debug_force_bn(bn); L.pd_debug_force_bres(1)
            n0 = L.pd_debug_bres_launches()
            try:
                us = run(*s_, iters=a.iters)[0]
            except RuntimeError:
                us = None
            took = L.pd_debug_bres_launches() > n0
            L.pd_debug_force_cta_group(0); L.pd_debug_force_bn(0); L.pd_debug_force_bres(0)
            row.append(us if took else None)
        print("%4d %3dx%-3d %5d %5d %2d %3d | %9.1f | " % (*s_, base) + " | ".join("%9.1f" % r if r is not None else "        -" for r in row))
elif a.small:
    print("   B   HxW     C     N ks res |     us")
    for s_ in [(1, 1, 128, 64, 64, 1, 0), (1, 1, 128 * 148, 64, 64, 1, 0), (1, 1, 128 * 148, 320, 128, 1, 0), (1, 1, 128 * 148, 320, 256, 1, 0),
               (1, 1, 128 * 148 * 2, 320, 256, 1, 0), (1, 1, 128 * 148 * 4, 320, 256, 1, 0), (1, 1, 128 * 148 * 4, 320, 256, 1, 1),
               (1, 1, 128 * 148, 1280, 256, 1, 0), (1, 1, 128 * 148, 2560, 256, 1, 0), (1, 1, 128 * 148, 5120, 256, 1, 0)]:
        us, tf, gb = run(*s_, iters=a.iters)
        print("%4d %3dx%-6d %5d %5d %2d %3d | %7.1f" % (*s_, us))
elif a.modes:
    from prompt_diffusion_b200 import _lib
    print("   B   HxW     C     N ks res cg |  full us | noMMA us | noTMA us")
    for s_ in [(16, 64, 64, 320, 320, 3, 0), (16, 64, 64, 320, 320, 1, 0), (16, 32, 32, 1280, 1280, 3, 0), (16, 64, 64, 320, 2560, 1, 0),
               (16, 16, 16, 1280, 1280, 3, 1), (16, 8, 8, 1280, 1280, 3, 1)]:
        for cg in (1, 2):
            _lib.lib.pd_debug_force_cta_group(cg)
            row = []
            for mode in (0, 1, 2):
                _lib.lib.pd_debug_gemm_mode(mode)
                row.append(run(*s_, iters=a.iters)[0])
            _lib.lib.pd_debug_gemm_mode(0)
            print("%4d %3dx%-3d %5d %5d %2d %3d %2d | %8.1f | %8.1f | %8.1f" % (*s_, cg, *row))
    _lib.lib.pd_debug_force_cta_group(0)
elif os.environ.get("PD_HALF_A"):
    # what-if (debug library): every CTA fetches only half of its A tile's rows (mode 12) against the full kernel
    from prompt_diffusion_b200 import _lib
    print("   B   HxW     C     N ks res cg |  full us | half-A us | no-epilogue us | half-A + no-epilogue us")
    for s_ in [(16, 64, 64, 320, 320, 1, 0), (16, 64, 64, 320, 320, 1, 1), (16, 64, 64, 320, 960, 1, 0), (16, 64, 64, 320, 2560, 1, 0), (16, 64, 64, 320, 320, 3, 0),
               (16, 64, 64, 640, 320, 3, 0), (16, 32, 32, 640, 640, 1, 1), (16, 32, 32, 640, 640, 3, 0), (16, 16, 16, 1280, 1280, 3, 0), (16, 8, 8, 1280, 1280, 3, 1)]:
        for cg in (1, 2):
            _lib.lib.pd_debug_force_cta_group(cg)
            row = []
            for mode in (0, 12, 8):
                _lib.lib.pd_debug_gemm_mode(mode)
                row.append(run(*s_, iters=a.iters)[0])
            _lib.lib.pd_debug_gemm_mode(0)
            print("%4d %3dx%-3d %5d %5d %2d %3d %2d | %8.1f | %9.1f | %9.1f" % (*s_, cg, *row))
    _lib.lib.pd_debug_force_cta_group(0)
elif a.epi_modes:
    from prompt_diffusion_b200 import _lib
    print("   B   HxW     C     N ks res |  full | noMMA | noTMA | noStore | noEpi | noMMA+noTMA (us)")
    for s_ in [(16, 64, 64, 320, 320, 1, 0), (16, 64, 64, 320, 960, 1, 0), (16, 64, 64, 320, 2560, 1, 0), (16, 32, 32, 640, 640, 1, 0),
               (16, 16, 16, 1280, 1280, 1, 0), (16, 64, 64, 320, 320, 3, 0)]:
        _lib.lib.pd_debug_force_cta_group(1)
        row = []
        for mode in (0, 1, 2, 3, 8, 9):
            _lib.lib.pd_debug_gemm_mode(mode)
            row.append(run(*s_, iters=a.iters)[0])
        _lib.lib.pd_debug_gemm_mode(0)
        print("%4d %3dx%-3d %5d %5d %2d %3d | " % s_ + " | ".join("%6.1f" % r for r in row))
    _lib.lib.pd_debug_force_cta_group(0)
elif a.one:
    if os.environ.get("PD_BENCH_BRES"):          # "cg,bn": resident-B schedule with that tile (ncu captures)
        from prompt_diffusion_b200 import _lib
        cg_, bn_ = [int(v) for v in os.environ["PD_BENCH_BRES"].split(",")]
        _lib.lib.pd_debug_force_cta_group(cg_); _lib.lib.pd_debug_force_bn(bn_); _lib.lib.pd_debug_force_bres(1)
    B, H, W, C, N, ks, res = a.one
    us, tf, gb = run(B, H, W, C, N, ks, bool(res), iters=a.iters, warm=2)
    print(f"B{B} {H}x{W} C{C}->N{N} k{ks} res={res}: {us:.1f} us  {tf:.0f} TFLOP/s  {gb:.0f} GB/s")
else:
    shapes = [  # (B,H,W,C,N,ks,res)
        (16, 64, 64, 320, 320, 1, 0), (16, 64, 64, 320, 320, 1, 1), (16, 64, 64, 320, 960, 1, 0),
        (16, 64, 64, 320, 2560, 1, 0), (16, 64, 64, 1280, 320, 1, 1), (16, 64, 64, 320, 320, 3, 0),
        (16, 64, 64, 320, 320, 3, 1), (16, 64, 64, 640, 640, 3, 0), (16, 32, 32, 640, 640, 1, 1),
        (16, 32, 32, 640, 5120, 1, 0), (16, 32, 32, 640, 640, 3, 1), (16, 32, 32, 1280, 1280, 3, 0),
        (16, 16, 16, 1280, 1280, 1, 1), (16, 16, 16, 1280, 10240, 1, 0), (16, 16, 16, 1280, 1280, 3, 1),
        (16, 16, 16, 2560, 1280, 3, 0), (16, 8, 8, 1280, 1280, 3, 1), (16, 8, 8, 2560, 1280, 3, 0),
        (16, 8, 8, 1280, 1280, 1, 0),
    ]
    from prompt_diffusion_b200 import _lib
    print("   B   HxW     C     N ks res |  auto us  TFLOP/s |  cg1 us  TFLOP/s |  cg2 us  TFLOP/s | cg1+sk us TFLOP/s | cg2+sk us TFLOP/s")
    for s in shapes:
        row = []
        for cg, sk in ((0, 0), (1, 0), (2, 0), (1, 1), (2, 1)):
            _lib.lib.pd_debug_force_cta_group(cg); _lib.lib.pd_debug_force_stream_k(sk)
            us, tf, gb = run(*s, iters=a.iters)
            row += [us, tf]
        _lib.lib.pd_debug_force_cta_group(0); _lib.lib.pd_debug_force_stream_k(0)
        print("%4d %3dx%-3d %5d %5d %2d %3d | %8.1f %8.0f | %7.1f %8.0f | %7.1f %8.0f | %8.1f %8.0f | %8.1f %8.0f" % (s[0], s[1], s[2], s[3], s[4], s[5], s[6], *row))
