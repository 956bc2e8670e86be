import os, sys, torch
sys.path.insert(0, "/root/repo")
from recommendflow_b200.dense_ops import sdpa
torch.manual_seed(0)
NB,S,dh=4,50,64
q=torch.randn(NB,S,dh,device="cuda"); k=torch.randn(NB,S,dh,device="cuda"); v=torch.randn(NB,S,dh,device="cuda")
o=sdpa(q,k,v,None,precision="tf32"); torch.cuda.synchronize()
mode=os.environ.get("RF_SDPA_DEBUG","0")
logits=(q@k.transpose(1,2))
if mode=="2":
    print("raw logits row0:", o[0,0,:6].tolist(), "ref", logits[0,0,:6].tolist()); print("seq1 row3:", o[1,3,:4].tolist(), "ref", logits[1,3,:4].tolist()); print("max err", float((o[:,:,:50]-logits[:,:,:50]).abs().max()))
elif mode=="1":
    p=torch.softmax(logits/8,dim=-1); print("P row0:", o[0,0,:6].tolist(), "ref", p[0,0,:6].tolist()); print("max err", float((o[:,:,:50]-p).abs().max()), "row sums", o[0,0].sum().item(), o[3,49].sum().item())
elif mode=="3":
    p=torch.softmax(logits/8,dim=-1); pp=p[0]@p[0].T; print("P.P^T row0:", o[0,0,:5].tolist(), "ref", pp[0,:5].tolist())
elif mode=="4":
    vp=torch.cat([v[0], v[1][:14]],0)   # keys 0..63 of the pair tile = seq0 (50) + first 14 rows of seq1
    ref=q[0]@vp[:64]; print("Q.V64 row0:", o[0,0,:5].tolist(), "ref", ref[0,:5].tolist(), "maxerr", float((o[0]-ref).abs().max()))
else:
    ref=sdpa(q,k,v,None,precision="fp32"); print("out row0:", o[0,0,:4].tolist(), "ref", ref[0,0,:4].tolist(), "max err", float((o-ref).abs().max()))
