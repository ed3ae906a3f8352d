"""N = 16384 (twice the largest benchmark size): factorisation, predictions at training inputs, and the analytic
gradient against a central difference of the log-likelihood along a random direction.  python tools/big_n_check.py"""
import os, sys, time, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from andvaranaut_b200.gp import GPEngine
N,d=16384,6
rng=np.random.default_rng(1)
X=rng.uniform(size=(N,d)); y=np.sin(X@np.linspace(0.5,2,d))+0.01*rng.normal(size=N); y=(y-y.mean())/y.std()
eng=GPEngine(nx=d,kerns=['Matern52'],noise=True,device='cuda:0'); eng.set_data(X,y)
th=np.r_[1e-3,np.ones(d),1.5]
t0=time.time(); info=eng.factorize(th); torch.cuda.synchronize(); print('factorize s',time.time()-t0,'info',int(info[0]))
idx=rng.choice(N,300,replace=False)
mu,var=eng.predict(X[idx]); mu=mu.cpu().numpy(); var=var.cpu().numpy()
print('train-point residual max',np.max(np.abs(mu-y[idx])),'var range',var.min(),var.max())
assert np.all(var>=1e-3*(1-1e-6)) and np.all(var<=1.5+1e-3+1e-9)
ll,g,info=eng.loglik_grad(th[None,:]); torch.cuda.synchronize()
t0=time.time(); ll,g,info=eng.loglik_grad(th[None,:]); torch.cuda.synchronize(); print('loglik_grad s',time.time()-t0)
ll=float(ll[0]); g=g[0].cpu().numpy()
dirv=rng.normal(size=th.shape)*th*1e-4
lp=float(eng.loglik_grad((th+dirv)[None,:],want_grad=False)[0][0]); lm=float(eng.loglik_grad((th-dirv)[None,:],want_grad=False)[0][0])
fd=(lp-lm)/2; an=float(g@dirv)
print('ll',ll,'directional derivative: analytic',an,'central difference',fd,'rel',abs(an-fd)/abs(fd))
assert abs(an-fd)<=1e-5*abs(fd)
print('N=16384 ok')
