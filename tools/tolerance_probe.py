import os, sys
import numpy as np
ROOT='/root/repo'
sys.path.insert(0, ROOT); sys.path.insert(0, ROOT+'/tests'); sys.path.insert(0, ROOT+'/tests/golden')
import torch
from oracle import gp_oracle as go
import cases
import test_gpu_parity as tp
import test_oracle as to
def grad_err(g, ref): return float(np.max(np.abs(g-ref)/np.maximum(np.abs(ref), 1e-3*np.max(np.abs(ref)))))
for name,(spec,N) in tp.SPECS.items():
    if 'Exponential' in spec.kerns and len(spec.kerns) > 2:
        X,y,th,_=cases.synth(spec,N,seed=17)
        rng=np.random.default_rng(4)
        thetas=np.stack([th, th*np.exp(0.05*rng.normal(size=th.shape))])
        eng=tp.engine(spec); eng.set_data(X,y)
        ll,grad,info=eng.loglik_grad(thetas)
        for b in range(2):
            r=go.loglik(spec,thetas[b],X,y)
            print('loglik', name, b, 'll rel', abs(float(ll[b])-r.ll)/abs(r.ll), 'grad', grad_err(grad[b].cpu().numpy(), r.grad))
for name,N,M in [('expo',90,3),('mix3',100,10)]:
    spec=to.SPECS[name]
    X,y,th,Xs=cases.synth(spec,N,seed=31,M=M)
    eng=tp.engine(spec); eng.set_data(X,y); eng.factorize(th)
    for pn in (True,False):
        rm,rv,rdm,rdv=go.predict_grad(spec,th,X,y,Xs,pred_noise=pn)
        m,v,dm,dv=(t.cpu().numpy() for t in eng.predict_grad(Xs,pred_noise=pn))
        kv=float(np.max(go.unpack(spec,th)['kv']))
        print('predict_grad', name, pn, 'dvar err', np.max(np.abs(dv-rdv))/max(np.max(np.abs(rdv)),kv), 'dmu', np.max(np.abs(dm-rdm))/np.max(np.abs(rdm)))
