import sys,os; sys.path.insert(0,'/root/repo')
from fmwr_b200 import _lib as L
ctx=L.Context(0)
n,F,k=2_000_000,39,32; field=1_000_000//F; p=field*F
d=L.Data.synth(ctx,n,[field]*F,None,0,1,0.1,20240601)
mc=L.ModelCfg(task=L.CLASSIFICATION,keep_w0=1,keep_w1=1,k=k,l1_w1=1e-3,l2_w1=1e-3,l2_v=1e-3)
m=L.Model(ctx,mc,p,L.F32); m.init_random(0,0.01,3)
it=60_000
sc=L.SolverCfg(solver=L.FTRL,max_iter=it,random_step=1,learn_rate=0.01,alpha_w=0.1,alpha_v=0.1,beta_w=1,beta_v=1,gamma=1e-4,min_target=-1,max_target=1,mode=L.MODE_EXACT,batch_size=1,precision=L.F32,compat=L.COMPAT_REFERENCE,step_size=-1)
ctx.timer_start(); L.train_dev(ctx,m,d,sc); ms=ctx.timer_stop_ms()
print('ftrl exact %.2f us per sample'%(ms*1e3/it))
