import sys, numpy as np, torch
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from andvaranaut_b200.gp import GPEngine
import bench
_, X, y, _ = bench.workload_c2()
y=(np.log(y)-np.log(y).mean())/np.log(y).std()
d=8
for kerns,ops in ((['Matern52'],[]),(['RBF','Matern52'],['+']),(['RBF','Matern52'],['*'])):
    eng=GPEngine(nx=d,kerns=kerns,ops=ops,noise=True,device='cuda:0'); eng.set_data(X,y)
    nk=len(kerns)
    th=np.r_[1e-4,0.7*np.ones(d*nk),1.5*np.ones(nk)]
    ths=torch.as_tensor(bench.theta_cloud(th,64,seed=1),device='cuda:0')
    for _ in range(3): eng.loglik_grad(ths)
    eng.set_profiling(True); eng.loglik_grad(ths); pm=eng.phase_ms()
    print(kerns,ops,{k:round(v,3) for k,v in pm.items() if v>0})
