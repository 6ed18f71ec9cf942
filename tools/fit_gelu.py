"""Coefficients of the one-MUFU GELU / GELU' of csrc/gemm.cu (gelu_q2, gelu_bwd2).

Phi(-t) and gelu'(-t) = Phi(-t) - t phi(t) are written as exp(-t^2/2 - b t) * polynomial(t) on [0, 6.5]; the polynomial is
fitted for minimal absolute error of the PRODUCT (Lawson-reweighted least squares), b is scanned, and the result is
re-evaluated in float32 Horner arithmetic.  CPU only (numpy + scipy); prints the coefficients pasted into the kernel.
"""
import numpy as np
from scipy.special import erfc
T=6.5
t=np.linspace(0,T,26001)
Phi_neg=0.5*erfc(t/np.sqrt(2))
phi=np.exp(-t*t/2)/np.sqrt(2*np.pi)
D=Phi_neg-t*phi
def fit(target, E, deg, iters=200):
    P=target/E
    V=np.vander(t/T,deg+1,increasing=True)
    w=np.ones_like(t); best=None
    for it in range(iters):
        W=np.sqrt(w)*E
        c,*_=np.linalg.lstsq(V*W[:,None],P*W,rcond=None)
        err=np.abs(E*(V@c-P)); m=err.max()
        if best is None or m<best[0]: best=(m,c.copy())
        w=w*(err/m+1e-3); w/=w.sum()/len(w)
    m,c=best
    return m,c/(T**np.arange(deg+1))
def f32eval(c,b,tt):
    tt=tt.astype(np.float32); cc=c.astype(np.float32)
    L=np.float32(1.4426950408889634)
    a2=np.float32(-0.5*1.4426950408889634); a1=np.float32(-b*1.4426950408889634)
    arg=(tt*(tt*a2+a1)).astype(np.float32)
    e=np.exp2(arg.astype(np.float64)).astype(np.float32)
    acc=np.full_like(tt,cc[-1])
    for k in range(len(cc)-2,-1,-1): acc=(acc*tt+cc[k]).astype(np.float32)
    return (acc*e).astype(np.float32)
for name,target,deg,bs in (('cdf',Phi_neg,6,np.linspace(0.9,1.02,13)),('cdf',Phi_neg,5,np.linspace(0.7,1.0,16)),('cdf',Phi_neg,4,np.linspace(0.75,0.95,11)),('dgelu',D,5,np.linspace(0.4,0.6,11)),('dgelu',D,7,np.linspace(0.55,0.75,11))):
    res=[]
    for b in bs:
        E=np.exp(-t*t/2-b*t)
        m,c=fit(target,E,deg)
        res.append((m,b,c))
    m,b,c=min(res,key=lambda r:r[0])
    q=f32eval(c,b,t)
    print(name,deg,'b=%.4f'%b,'minimax',m,'fp32 err',np.abs(q-target).max())
    print('   coeffs', ', '.join('%.10ef'%x for x in c))
