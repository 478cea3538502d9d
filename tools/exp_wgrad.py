import sys, itertools
sys.path.insert(0, '.')
import torch
from ever_b200 import ops
def timeit(fn, iters=20, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3
for (n,h,w,cin,cout,k,s) in [(8,128,128,256,256,3,1),(8,64,64,256,256,3,1),(8,32,32,256,256,3,1),(8,16,16,512,512,3,1),(8,128,128,64,64,3,1),(8,32,32,1024,256,1,1)]:
    x = torch.randn(n,h,w,cin,device='cuda').bfloat16(); dy = torch.randn(n,h//s,w//s,cout,device='cuda').bfloat16()
    dw = torch.empty(cout,cin,k,k,device='cuda'); ws = torch.empty(256<<20, device='cuda')
    flop = 2.0*n*(h//s)*(w//s)*cout*cin*k*k
    res=[]
    for nt in (256,128,64):
        if nt > cout: continue
        for sp in (0,1,2,4,8,16,32):
            try:
                t = timeit(lambda: ops.conv2d_wgrad(x,dy,k,s,dw=dw,ws=ws,force_nt=nt,force_split=sp))
                res.append((t,nt,sp))
            except Exception as e:
                pass
    res.sort()
    print((n,h,w,cin,cout,k,s), ' | '.join('nt%d sp%d %.1fus %.0fTF'%(nt,sp,t,flop/t/1e6) for t,nt,sp in res[:6]), ' || default:', [('%.1f'%t) for t,nt,sp in res if sp==0])
