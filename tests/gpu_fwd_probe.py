import subprocess, sys, os
CASES = [(4,8,63,63,64,1,"float16"),(4,8,63,63,64,0,"float16"),(1,1,128,128,64,0,"bfloat16"),(1,1,256,256,64,0,"bfloat16"),
         (1,2,256,256,64,1,"bfloat16"),(1,1,63,63,64,1,"bfloat16"),(1,1,300,300,64,1,"bfloat16"),(4,8,127,127,128,1,"float16"),(2,4,1024,1024,64,1,"bfloat16"), (8,16,1024,1024,64,1,"bfloat16"),(1,2,1024,256,64,1,"bfloat16"),(2,3,1000,333,128,1,"float16"),(1,1,64,1024,64,1,"bfloat16")]
child = r'''
import sys, torch
sys.path.insert(0, ".")
from flash_cosine_sim_attention_b200.flash_cosine_sim_attention import _ext
B,H,Nq,Nk,D,causal,dt = sys.argv[1:8]
B,H,Nq,Nk,D,causal = map(int,(B,H,Nq,Nk,D,causal)); dt = getattr(torch, dt)
q = torch.nn.functional.normalize(torch.randn(B,H,Nq,D,device="cuda"),dim=-1).to(dt); k = torch.nn.functional.normalize(torch.randn(B,H,Nk,D,device="cuda"),dim=-1).to(dt); v = torch.randn(B,H,Nk,D,device="cuda").to(dt)
o, inv_l = _ext().forward_ex(q,k,v,None,None,False,None,8.0,8.0,bool(causal),0,True,False)[:2]
torch.cuda.synchronize()
s = torch.einsum("bhid,bhjd->bhij", q.float(), k.float())*8-8
if causal: s = s.masked_fill(torch.ones(Nq,Nk,dtype=torch.bool,device="cuda").triu(Nk-Nq+1), float("-inf"))
p = s.exp(); ref = (p@v.float())/p.sum(-1,keepdim=True)
print("max err", (o.float()-ref).abs().max().item())
'''
for c in CASES:
    try:
        r = subprocess.run([sys.executable, "-c", child] + [str(x) for x in c], capture_output=True, text=True, timeout=25 if c is not CASES[0] else 120)
        print(c, "rc", r.returncode, r.stdout.strip()[-60:], r.stderr.strip()[-300:] if r.returncode else "", flush=True)
    except subprocess.TimeoutExpired:
        print(c, "HANG", flush=True)
