import os, sys, math
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from prompt_diffusion_b200 import ops, _lib
from prompt_diffusion_b200._lib import PD_ENGINE_TC
B,H,W,C,N,ks,res = [int(v) for v in sys.argv[1:8]]
if len(sys.argv) > 8: _lib.lib.pd_debug_force_cta_group(int(sys.argv[8]))
if len(sys.argv) > 9: _lib.lib.pd_debug_force_bn(int(sys.argv[9]))
if len(sys.argv) > 10: _lib.lib.pd_debug_force_bres(int(sys.argv[10]))
if len(sys.argv) > 11: _lib.lib.pd_debug_force_stream_k(int(sys.argv[11]))
dev="cuda"; M=B*H*W
x=torch.randn(M,C,device=dev).to(torch.bfloat16); w=(torch.randn(N,ks*ks*C,device=dev)/math.sqrt(ks*ks*C)).to(torch.bfloat16)
bias=torch.randn(N,device=dev); out=torch.empty(M,N,device=dev,dtype=torch.bfloat16)
r=torch.randn(M,N,device=dev).to(torch.bfloat16) if res else None
for _ in range(3): ops.conv2d(x,w,out,B,H,W,ksize=ks,bias=bias,res=r,engine=PD_ENGINE_TC)
dbg=torch.zeros(5*64*2,dtype=torch.int64,device=dev)
_lib.lib.pd_debug_timeline(dbg.data_ptr())
ops.conv2d(x,w,out,B,H,W,ksize=ks,bias=bias,res=r,engine=PD_ENGINE_TC)
torch.cuda.synchronize(); _lib.lib.pd_debug_timeline(None)
d=dbg.cpu().reshape(5,64,2); t0=int(d[d>0].min())
print(f"shape B{B} {H}x{W} C{C} N{N} k{ks} res{res}; times in us since first stamp (CTA 0)")
print("tile | TMA start  lastwait | MMA start   done | EPI start   done | acc in regs  staging ready | math done  fenced+barrier")
for i in range(64):
    if d[0,i,0]==0 and d[1,i,0]==0: break
    f=lambda v: f"{(int(v)-t0)/1e3:8.2f}" if v>0 else "     -  "
    print(f"{i:4d} | {f(d[0,i,0])} {f(d[0,i,1])} | {f(d[1,i,0])} {f(d[1,i,1])} | {f(d[2,i,0])} {f(d[2,i,1])} | {f(d[3,i,0])} {f(d[3,i,1])} | {f(d[4,i,0])} {f(d[4,i,1])}")
