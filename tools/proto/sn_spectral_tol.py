import numpy as np, sys
exec(open(''+__import__('os').path.dirname(__import__('os').path.abspath(__file__))+'/sn_spectral_proto.py').read().split("alo=az.min()")[0])
alo=az.min()
rng=np.random.default_rng(7)
h=1-az
P=[]
for t in range(6000):
    mode=t%3
    Om=rng.uniform(0.0,1.2); w0=rng.uniform(-3.5,0.5)
    if mode==0: OK=0;Ode=1-Om;w1=0
    elif mode==1: Ode=rng.uniform(0,1.6); OK=1-Om-Ode; w1=0; w0=-1
    else: Om=rng.uniform(0.05,1);Ode=rng.uniform(0,1.5); OK=1-Om-Ode; w0=rng.uniform(-3,0); w1=rng.uniform(-3,2)
    P.append((mode,Om,OK,Ode,w0,w1))
EX={}
for p in P:
    R,D=exact(*p[1:]); EX[p]=(R,D)
for M in (24,28,32):
    ac,C,W,Dm=build(M,alo)
    dmax=np.abs(Dm/h[:,None]).max(axis=0)
    res=[]
    for p in P:
        qv=Q(ac,*p[1:])
        R,D=EX[p]
        if (qv<=0).any() or not np.isfinite(R).all(): continue
        q=1/np.sqrt(qv); c=C@q
        e=np.abs((W@c)/R-1).max()
        tail=np.abs(c[-3:]).max()/abs(c[0])
        tail2=np.abs(c[-2:]).max()/abs(c[0])
        B=(dmax*np.abs(c)).sum()/q.min()
        res.append((p[0],e,tail,B,np.abs(D/R).max(),tail2))
    res=np.array(res)
    for mode in (0,1,2):
        s=res[res[:,0]==mode]
        for tol in (1e-13,1e-12,1e-11,1e-10):
            ok=(s[:,2]<tol)
            okc=ok&(s[:,3]<0.5e-6)
            print("M=%d mode=%d tol=%.0e pass=%.3f maxerr=%.2e | with conv bound pass=%.3f (actual conv frac %.3f)"%(M,mode,tol,ok.mean(),s[ok,1].max(),okc.mean(),(s[:,4]<1e-6).mean()))
