import numpy as np, sys
sys.path.insert(0,'/root/repo')
exec(open(''+__import__('os').path.dirname(__import__('os').path.abspath(__file__))+'/sn_spectral_proto.py').read().split("alo=az.min()")[0])
alo=az.min()
from cosmopmc_b200 import targets as T
w,mean,cov=T.proposal_sn(10)
rng=np.random.default_rng(5)
N=4000
k=rng.integers(0,10,N)
X=np.array([rng.multivariate_normal(mean[i],cov[i]) for i in k])
lo,hi=T.target_sn_demo().box
inbox=((X>=lo)&(X<=hi)).all(axis=1)
X=X[inbox]; print("in box",len(X))
h=1-az
for M in (20,24,28,32,40):
    ac,C,W,Dm=build(M,alo)
    dmax=np.abs(Dm/h[:,None]).max(axis=0)
    res=[]
    for x in X:
        Om,w0=x[0],x[1]
        qv=Q(ac,Om,0,1-Om,w0,0)
        if (qv<=0).any(): res.append((1,1,1,0)); continue
        q=1/np.sqrt(qv)
        c=C@q
        R,D=exact(Om,0,1-Om,w0,0)
        Rc=W@c
        e=np.abs(Rc/R-1).max()
        tail=np.abs(c[-3:]).max()/abs(c[0])
        B=(dmax*np.abs(c)).sum()/q.min()
        actual=np.abs(D/R).max()
        res.append((e,tail,B,actual))
    res=np.array(res)
    for tol in (1e-13,3e-13,1e-12):
        ok=(res[:,1]<tol)&(res[:,2]<0.5e-6)
        print("M=%d tol=%.0e pass=%.4f maxerr(pass)=%.2e  convfail=%.4f actual_unconv=%.5f"%(M,tol,ok.mean(),res[ok,0].max(),(res[:,2]>=0.5e-6).mean(),(res[:,3]>1e-6).mean()))
