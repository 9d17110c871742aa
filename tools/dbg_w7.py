import sys, os, torch, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from clover_b200 import ops, swin, tables
from oracle import clover_oracle as O
torch.manual_seed(0)
dims=(4,14,14); shifted=int(sys.argv[1]) if len(sys.argv)>1 else 0; Bc=2; heads=3; hd=32
win,sh=O.get_window_size(dims,(8,7,7),(4,3,3) if shifted else (0,0,0))
N=win[0]*win[1]*win[2]; nwin=(dims[0]//win[0])*(dims[1]//win[1])*(dims[2]//win[2]); batch=Bc*nwin
g=torch.Generator().manual_seed(30)
qkv=(torch.randn(batch*N,3*heads*hd,generator=g)*0.7).bfloat16().cuda()
dout=torch.randn(batch*N,heads*hd,generator=g).bfloat16().cuda()
table=(torch.randn(2535,heads,generator=g)*0.5).cuda()
code,off=tables.rel_code(N,(8,7,7)); code=torch.from_numpy(code).cuda()
masked=any(s>0 for s in sh)
region=torch.from_numpy(tables.region_ids(*dims,win,sh)).cuda() if masked else None
spec=swin._w7_spec(dims,win,sh,(8,7,7),"cuda")
out=torch.empty(batch*N,heads*hd,dtype=torch.bfloat16,device="cuda"); lse=torch.empty(batch,heads,N,device="cuda")
kw=dict(bias_table=table,rel_code=code,code_off=off,region=region)
ops.attention_fwd(qkv,batch,N,heads,hd,out,lse,w7=spec,**kw)
d1=torch.empty_like(qkv); d2=torch.empty_like(qkv)
t1=torch.zeros(2535,heads,device="cuda"); t2=torch.zeros(2535,heads,device="cuda")
ops.attention_bwd(qkv,out,dout,lse,batch,N,heads,hd,d1,1.0,dbias_table=t1,w7=spec,**kw)
ops.attention_bwd(qkv,out,dout,lse,batch,N,heads,hd,d2,1.0,dbias_table=t2,**kw)
torch.cuda.synchronize()
a=d1.float().view(batch,N,3,heads,hd); b=d2.float().view(batch,N,3,heads,hd)
for part,name in enumerate("qkv"):
    for h in range(heads):
        e=(a[:,:,part,h]-b[:,:,part,h]).norm(dim=-1)/ (b[:,:,part,h].norm(dim=-1)+1e-6)   # [batch,N]
        bad=(e>0.05)
        print("d"+name,"head",h,"rel", float((a[:,:,part,h]-b[:,:,part,h]).norm()/b[:,:,part,h].norm()), "bad rows", int(bad.sum()), "of", bad.numel(),
              "first bad", bad.nonzero()[:6].tolist())
print("dtab rel", float((t1-t2).norm()/t2.norm()))
