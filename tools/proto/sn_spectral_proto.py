import numpy as np, sys
np.seterr(all='ignore')
zs=[]
for l in open('/root/repo/tests/golden/sn_union307.txt'):
    if l[0] in '#@': continue
    zs.append(float(l.split()[0]))
zs=np.unique(np.array(zs)); nz=len(zs)
az=1/(1+zs)
# Romberg stage-5 nodes, weights: g0=(f(a)+f(1))/2, S1: mid; S2: 2 nodes; S3: 4; S4: 8
ROMBW=[3937.0/103275.0,3062.0/80325.0,27728.0/722925.0,22016.0/722925.0,65536.0/722925.0]
ROMBD=[-31.0/206550.0,-73.0/481950.0,-67.0/722925.0,-424.0/722925.0,256.0/722925.0]
def nodes(a):
    h=1-a
    t=[0.0,1.0]; w=[0.5*ROMBW[0]]*2; d=[0.5*ROMBD[0]]*2
    for j in range(1,5):
        it=1<<(j-1); 
        for i in range(it):
            t.append((i+0.5)/it); w.append(ROMBW[j]); d.append(ROMBD[j])
    t=np.array(t); return a+h*t, h*np.array(w), h*np.array(d)
def Q(a,Om,OK,Ode,w0,w1):
    return Om+OK*a+Ode*a**(-3*(w0+w1))*np.exp(-3*w1*(1-a))
def exact(Om,OK,Ode,w0,w1):
    R=np.zeros(nz);D=np.zeros(nz)
    for k,a in enumerate(az):
        x,w,d=nodes(a)
        f=x**-0.5/np.sqrt(Q(x,Om,OK,Ode,w0,w1))
        R[k]=np.dot(w,f); D[k]=np.dot(d,f)
    return R,D
def build(M,alo):
    # Chebyshev nodes first kind on [alo,1]
    j=np.arange(M); xc=np.cos(np.pi*(j+0.5)/M); ac=0.5*(1+alo)+0.5*(1-alo)*xc
    # coefficient matrix: c_m = (2/M) sum_j f_j cos(m pi (j+.5)/M), c_0 halved
    C=(2.0/M)*np.cos(np.pi*np.outer(np.arange(M),(j+0.5))/M); C[0]*=0.5
    W=np.zeros((nz,M));Dm=np.zeros((nz,M))
    for k,a in enumerate(az):
        x,w,d=nodes(a)
        xx=(2*x-(1+alo))/(1-alo)
        Tm=np.cos(np.outer(np.arange(M),np.arccos(np.clip(xx,-1,1))))  # M x 17
        W[k]=Tm@(w*x**-0.5); Dm[k]=Tm@(d*x**-0.5)
    return ac,C,W,Dm
alo=az.min()
rng=np.random.default_rng(1)
for M in (16,20,24,28,32):
    ac,C,W,Dm=build(M,alo)
    WN=W@C; 
    worst=0;worstd=0;wp=None
    for trial in range(400):
        Om=rng.uniform(0.0,1.2); w0=rng.uniform(-3.5,0.5); 
        if trial%2: OK=0; Ode=1-Om; w1=0
        else: Ode=rng.uniform(0,1.5); OK=1-Om-Ode; w1=rng.uniform(-2,2)
        qv=Q(ac,Om,OK,Ode,w0,w1)
        if (qv<=0).any(): continue
        f=1/np.sqrt(qv)
        R,D=exact(Om,OK,Ode,w0,w1)
        if not np.isfinite(R).all(): continue
        Rc=WN@f
        c=C@f
        tail=np.abs(c[-3:]).max()/np.abs(c[0])
        e=np.abs(Rc/R-1).max()
        if e>worst: worst=e; wp=(Om,OK,Ode,w0,w1,tail)
    print(M,worst,wp)
print("---- distribution")
for M in (20,24,28,32):
    ac,C,W,Dm=build(M,alo); WN=W@C
    errs=[];tails=[]
    for trial in range(3000):
        Om=rng.uniform(0.0,1.2); w0=rng.uniform(-3.5,0.5)
        mode=trial%3
        if mode==0: OK=0; Ode=1-Om; w1=0
        elif mode==1: Ode=rng.uniform(0,1.5); OK=1-Om-Ode; w1=rng.uniform(-2,2)
        else:
            Om=rng.normal(0.3,0.12); w0=rng.normal(-1.0,0.4); OK=0;Ode=1-Om;w1=0
        qv=Q(ac,Om,OK,Ode,w0,w1)
        if (qv<=0).any(): continue
        f=1/np.sqrt(qv)
        R,D=exact(Om,OK,Ode,w0,w1)
        if not np.isfinite(R).all(): continue
        c=C@f
        tail=np.abs(c[-4:]).max()/np.abs(c[0])
        e=np.abs(WN@f/R-1).max()
        errs.append((mode,e,tail))
    errs=np.array(errs)
    for mode in (0,1,2):
        s=errs[errs[:,0]==mode]
        ok=s[:,2]<1e-14
        print(M,mode,len(s),"frac tail<1e-14:",ok.mean(),"max err among ok:",s[ok,1].max() if ok.any() else None, "frac err<1e-12:",(s[:,1]<1e-12).mean())
