import os, sys, torch
ROOT="/root/repo"
for p in (ROOT, ROOT+"/gif-gan_b200"): sys.path.insert(0,p)
from collections import OrderedDict
from gifgan import ops, _cabi
ops.set_precision("bf16", tensor_cores=True)
st = ops.reset_default_store(device="cuda")
import numpy as np
B,H,K=64,64,64
wv = st.get_variable("w", (5,5,3,K), lambda r,s: np.zeros(s,dtype="float32")+0.01, filter_taps=25)
st.finalize(OrderedDict(all=[wv]))
g = ops._Geom(B,(1,H,H),3,(1,H//2,H//2),K,(1,5,5),(1,2,2),(0,1,1))
large = torch.randn(B,H,H,3,device="cuda")
small = torch.empty(B,H//2,H//2,K,device="cuda",dtype=torch.bfloat16)
b = ops._BnInfo(); b.pre=torch.randn(B,H//2,H//2,K,device="cuda"); b.mean=torch.zeros(1,K,device="cuda"); b.rstd=torch.ones(1,K,device="cuda")
b.gamma=torch.ones(K,device="cuda"); b.beta=torch.zeros(K,device="cuda"); b.act,b.act_param,b.groups,b.Cc="relu",0.0,1,K
def t(fn,reps=20):
    fn(); torch.cuda.synchronize()
    _cabi.lib().gg_debug_set_repeat(reps)
    best=1e9
    for _ in range(3):
        s,e=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record(); e.synchronize(); best=min(best,s.elapsed_time(e)/reps)
    _cabi.lib().gg_debug_set_repeat(1)
    return best*1e3
print("c3m_down plain us", t(lambda: ops._run_down(g,large,wv,None,torch.bfloat16,None,0.0,4,out=small)))
with ops.stats_arena():
    print("c3m_down + bnb us", t(lambda: ops._run_down(g,large,wv,None,torch.bfloat16,None,0.0,4,out=small,bnb=b)))
