import sys, torch
sys.path.insert(0, ".")
from clover_b200 import ops
dev="cuda"
for rows, C, xdt in ((4096, 768, torch.bfloat16), (640, 768, torch.bfloat16), (4096, 768, torch.float32), (29184, 768, torch.bfloat16), (100352, 512, torch.float32)):
    x=torch.randn(rows,C,device=dev).to(xdt); g=torch.ones(C,device=dev); b=torch.zeros(C,device=dev)
    mean=torch.zeros(rows,device=dev); rstd=torch.ones(rows,device=dev)
    dy=torch.randn(rows,C,device=dev).to(torch.bfloat16)
    dx=torch.empty(rows,C,device=dev); dx16=torch.empty(rows,C,device=dev,dtype=torch.bfloat16)
    small=torch.zeros(2*C,device=dev)
    def f():
        if xdt==torch.bfloat16: ops.lnr_bwd(x,g,b,1e-5,mean,rstd,dy,dx_bf16=dx16,dgamma=small[:C],dbeta=small[C:])
        else: ops.lnr_bwd(x,g,b,1e-5,mean,rstd,dy,dx=dx,dgamma=small[:C],dbeta=small[C:])
    for _ in range(3): f()
    torch.cuda.synchronize()
    e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(50): f()
    e1.record(); torch.cuda.synchronize()
    print(rows,C,xdt, 'us per launch', e0.elapsed_time(e1)/50*1e3)
    def f2():
        ops.lnr_bwd(x,g,b,1e-5,mean,rstd,dy,dx_bf16=dx16) if xdt==torch.bfloat16 else ops.lnr_bwd(x,g,b,1e-5,mean,rstd,dy,dx=dx)
    for _ in range(3): f2()
    torch.cuda.synchronize()
    e0.record()
    for _ in range(50): f2()
    e1.record(); torch.cuda.synchronize()
    print('   without dgamma/dbeta:', e0.elapsed_time(e1)/50*1e3)
