import sys, numpy as np, torch
sys.path.insert(0,'/root/repo')
from tsr_b200 import ops, synth
rs=np.random.RandomState(5)
M=32
S=(rs.rand(2,296,36,100)<0.136).astype(np.uint8)
T,bgd=synth.templates(M=M,L=40,Fd=36,Kd=100)
W,c=synth.linear_filters(T,bgd)
fb=ops.pack_features(torch.from_numpy(np.ascontiguousarray(S.reshape(2,296,-1))).cuda())
Sf=torch.from_numpy(S.reshape(2,296,-1).astype(np.float64)).cuda()
Wd=torch.from_numpy(W.reshape(M,40,-1).astype(np.float64)).cuda()
ref=torch.zeros(2,296,M,dtype=torch.float64,device='cuda')
for t in range(296-40+1):
    ref[:,t,:]=torch.einsum('bld,mld->bm',Sf[:,t:t+40],Wd)
ref=ref.cpu().numpy()
scale=np.abs(ref).max()
for name,pl in (("fp16",0),("bf16",1),("int8x2",3),("int8x3",4)):
    got=ops.detect(fb,ops.FilterBank(W,planes=pl)).cpu().numpy().astype(np.float64)
    err=np.abs(got-ref)[:,:257]
    r=np.abs(ref)[:,:257]
    floor=np.maximum(err-1e-5*r,0).max()/scale
    print("%-7s max|err|/scale %.2e   rms %.2e  floor needed %.2e   max per-element rel (|want|>1e-3 scale) %.2e  scale %.1f min|want| %.3g"%(name,err.max()/scale,np.sqrt((err**2).mean())/scale,floor,(err/np.maximum(r,1e-30))[r>1e-3*scale].max(),scale,r.min()))
# reference float32 sequential
