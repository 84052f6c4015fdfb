import sys, ctypes, numpy as np, torch
sys.path.insert(0,'/root/repo')
from pymcz_b200 import _lib, _scales as sc, ops
from pymcz_b200.synth import synthetic_table
L=_lib.lib()
meas,err=synthetic_table(125000,config=4,start=0,stop=4000)
m=torch.from_numpy(meas).cuda(); e=torch.from_numpy(err).cuda()
out=(ctypes.c_ulonglong*8)()
L.mcz_debug_counters(out,1)
pct,nv,st,tr,ret,Z=ops.forward_production(m,e,2200,sc.T_ALL)
torch.cuda.synchronize()
L.mcz_debug_counters(out,1)
print('counters [total, fast, no_subsample, general, crowded, trivial, narrow_rounds]',list(out))
st=st.cpu().numpy(); nv=nv.cpu().numpy()
# per key: fraction of galaxies with status ok & nvalid
for j,k in enumerate(sc.KEYS):
    if (st[:,j]!=1).any(): print(k, 'ok%.3f empty%.3f nonpos%.3f meanN %.0f'%((st[:,j]==0).mean(),(st[:,j]==2).mean(),(st[:,j]==3).mean(), nv[:,j].mean()))
