"""Big-integer model of the BLS12-377 ate pairing: derives the constants of csrc/pairing_params_gen.h (through
tools/gen_pairing_params.py) and serves the tests as the reference the product's host
pairing (csrc/pairing.h) is compared with bit for bit.  A development / test script: nothing in the product imports it.

Parity unpinned against arkworks' GT values: only pairing EQUATIONS are ever checked, and those hold for any bilinear
non-degenerate pairing.  Tower: Fq2 = Fq[u]/(u^2+5), Fq12 = Fq2[w]/(w^6 - u), D-type twist y^2 = x^3 + 1/u -- the constants of
ark-bls12-377 0.3.0 (reference Cargo.lock:77); Miller loop over the BLS parameter x; plain (q^12-1)/r final exponentiation.
The G2 generator is DERIVED (smallest x = (k, 1) on the twist, cofactor-cleared), not arkworks' constant: the reference's
SRS is a throw-away test SRS (README.md:26) and no G2 element enters the proof bytes or the transcript.
"""
import random, sys
q = 0x1ae3a4617c510eac63b05c06ca1493b1a22d9f300f5138f1ef3622fba094800170b5d44300000008508c00000000001
r = 0x12ab655e9a2ca55660b44d1e5c37b00159aa76fed00000010a11800000000001
X = 0x8508c00000000001
assert r == X**4 - X**2 + 1
assert q == ((X - 1)**2 * r) // 3 + X and ((X-1)**2 * r) % 3 == 0
t = X + 1
assert (q + 1 - t) % r == 0
NR = -5 % q  # u^2 = -5
# Fq2 as tuples
def f2add(a,b): return ((a[0]+b[0])%q,(a[1]+b[1])%q)
def f2sub(a,b): return ((a[0]-b[0])%q,(a[1]-b[1])%q)
def f2mul(a,b): return ((a[0]*b[0]+NR*a[1]*b[1])%q,(a[0]*b[1]+a[1]*b[0])%q)
def f2neg(a): return ((-a[0])%q,(-a[1])%q)
def f2inv(a):
    n = (a[0]*a[0]-NR*a[1]*a[1])%q
    ni = pow(n,-1,q)
    return (a[0]*ni%q,(-a[1])*ni%q)
def f2pow(a,e):
    res=(1,0)
    while e:
        if e&1: res=f2mul(res,a)
        a=f2mul(a,a); e>>=1
    return res
F2Z=(0,0); F2ONE=(1,0)
XI=(0,1)  # w^6 = u
BT = f2inv(XI)  # b' = 1/xi for D-type twist (b=1)
assert BT == (0, 155198655607781456406391640216936120121836107652948796323930557600032281009004493664981332883744016074664192874906)  # ark-bls12-377 g2 COEFF_B
# twist group order: candidates
t2 = t*t - 2*q
import math
f2sq = (4*q*q - t2*t2)//3
f = math.isqrt(f2sq); assert f*f==f2sq
cands = [q*q+1-(t2+3*f)//2, q*q+1-(t2-3*f)//2, q*q+1-(-t2+3*f)//2, q*q+1-(-t2-3*f)//2, q*q+1-t2, q*q+1+t2]
# EC over Fq2 (twist), affine with None = inf
def e2add(P,Q):
    if P is None: return Q
    if Q is None: return P
    if P[0]==Q[0]:
        if f2add(P[1],Q[1])==F2Z: return None
        lam = f2mul(f2mul((3,0),f2mul(P[0],P[0])), f2inv(f2mul((2,0),P[1])))
    else:
        lam = f2mul(f2sub(Q[1],P[1]), f2inv(f2sub(Q[0],P[0])))
    x3 = f2sub(f2sub(f2mul(lam,lam),P[0]),Q[0])
    y3 = f2sub(f2mul(lam,f2sub(P[0],x3)),P[1])
    return (x3,y3)
def e2mul(P,k):
    R=None
    while k:
        if k&1: R=e2add(R,P)
        P=e2add(P,P); k>>=1
    return R
def f2sqrt(a):
    # Fq2 sqrt via norm trick: a = (x + y u)^2
    if a==F2Z: return F2Z
    # generic: Tonelli-ish using q^2 = 1 mod ..; use algorithm: a^((q^2+..)) not simple since q=1 mod 4. Use Cipolla-like via norm:
    n = (a[0]*a[0]-NR*a[1]*a[1])%q
    if pow(n,(q-1)//2,q)!=1: return None
    s = fqsqrt(n)
    inv2 = pow(2,-1,q)
    for sg in (s,(-s)%q):
        d = (a[0]+sg)*inv2%q
        if pow(d,(q-1)//2,q)==1 or d==0:
            x = fqsqrt(d)
            if x==0: continue
            y = a[1]*pow(2*x,-1,q)%q
            if f2mul((x,y),(x,y))==a: return (x,y)
    return None
def fqsqrt(a):
    a%=q
    if a==0: return 0
    if pow(a,(q-1)//2,q)!=1: return None
    s,tt=0,q-1
    while tt%2==0: s+=1; tt//=2
    z=2
    while pow(z,(q-1)//2,q)!=q-1: z+=1
    m,c,t_,R=s,pow(z,tt,q),pow(a,tt,q),pow(a,(tt+1)//2,q)
    while t_!=1:
        i,t2_=0,t_
        while t2_!=1: t2_=t2_*t2_%q; i+=1
        b=pow(c,1<<(m-i-1),q); m,c=i,b*b%q; t_,R=t_*c%q,R*b%q
    return R
def find_g2(order):
    cof = order//r
    xx=1
    while True:
        x=(xx,1)
        rhs=f2add(f2mul(f2mul(x,x),x),BT)
        y=f2sqrt(rhs)
        if y is not None:
            P=(x,y)
            G=e2mul(P,cof)
            if G is not None:
                assert e2mul(G,r) is None, "order"
                return G, xx
        xx+=1
for c in cands:
    if c%r==0:
        try:
            G2,xx=find_g2(c); ORDER=c; break
        except AssertionError as e:
            pass
# Fq12 = Fq2[w]/(w^6 - xi): list of 6 Fq2
def f12mul(a,b):
    res=[F2Z]*11
    for i in range(6):
        if a[i]==F2Z: continue
        for j in range(6):
            if b[j]==F2Z: continue
            res[i+j]=f2add(res[i+j],f2mul(a[i],b[j]))
    out=[res[k] for k in range(6)]
    for k in range(6,11):
        out[k-6]=f2add(out[k-6],f2mul(res[k],XI))
    return out
F12ONE=[F2ONE]+[F2Z]*5
def f12pow(a,e):
    res=F12ONE
    while e:
        if e&1: res=f12mul(res,a)
        a=f12mul(a,a); e>>=1
    return res
# G1 affine over Fq
def e1add(P,Q):
    if P is None: return Q
    if Q is None: return P
    if P[0]==Q[0]:
        if (P[1]+Q[1])%q==0: return None
        lam=3*P[0]*P[0]*pow(2*P[1],-1,q)%q
    else: lam=(Q[1]-P[1])*pow(Q[0]-P[0],-1,q)%q
    x3=(lam*lam-P[0]-Q[0])%q; y3=(lam*(P[0]-x3)-P[1])%q
    return (x3,y3)
def e1mul(P,k):
    R=None
    while k:
        if k&1: R=e1add(R,P)
        P=e1add(P,P); k>>=1
    return R
G1=(81937999373150964239938255573465948239988671502647976594219695644855304257327692006745978603320413799295628339695,241266749859715473739788878240585681733927191168601896383759122102112907357779751001206799952863815012735208165030)
assert (G1[1]**2-G1[0]**3-1)%q==0 and e1mul(G1,r) is None
def line(T,lam,P):
    # l = yP - lam' xP w + (lam' x_T - y_T) w^3
    c=[F2Z]*6
    c[0]=(P[1],0)
    c[1]=f2neg(f2mul(lam,(P[0],0)))
    c[3]=f2sub(f2mul(lam,T[0]),T[1])
    return c
def miller(P,Q):
    # f_{X,Q}(P), Q on twist
    if P is None or Q is None: return F12ONE
    f=F12ONE; T=Q
    bits=bin(X)[3:]
    for b in bits:
        lam=f2mul(f2mul((3,0),f2mul(T[0],T[0])),f2inv(f2mul((2,0),T[1])))
        f=f12mul(f12mul(f,f),line(T,lam,P))
        T=e2add(T,T)
        if b=='1':
            lam=f2mul(f2sub(Q[1],T[1]),f2inv(f2sub(Q[0],T[0])))
            f=f12mul(f,line(T,lam,P))
            T=e2add(T,Q)
    return f
FE=(q**12-1)//r
def pairing(P,Q): return f12pow(miller(P,Q),FE)
if __name__=='__main__':
    e=pairing(G1,G2)
    assert e!=F12ONE
    assert f12pow(e,r)==F12ONE
    a,b=random.randrange(1,r),random.randrange(1,r)
    e2=pairing(e1mul(G1,a),e2mul(G2,b))
    assert e2==f12pow(e,a*b%r), "bilinear"
    print('pairing OK')
