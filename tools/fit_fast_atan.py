"""Fit and error check of crossview_core.h::fast_atan2f (development aid): a degree-7 polynomial in t^2 for atan(t)/t on
[0, 1] by iteratively re-weighted least squares, then the float32 evaluation (emulated with numpy, ratio perturbed by
+-2 ulp like __fdividef) against float64 atan2 over 2.4e7 points.  Prints the coefficients and the worst error."""
import numpy as np
from numpy.polynomial import chebyshev as C
# fit f(u) = atan(sqrt(u))/sqrt(u) on u in [0,1] with degree-n polynomial in u (odd polynomial in t of degree 2n+1)
def fit(n):
    k = np.arange(4000)
    u = 0.5 - 0.5*np.cos(np.pi*(k+0.5)/4000)
    t = np.sqrt(u)
    f = np.where(t>0, np.arctan(t)/np.maximum(t,1e-300), 1.0)
    # minimax-ish via Remez-lite: iteratively reweighted LSQ
    w = np.ones_like(u)
    V = np.vander(u, n+1, increasing=True)
    for it in range(60):
        c = np.linalg.lstsq(V*w[:,None], f*w, rcond=None)[0]
        e = np.abs(V@c - f)*t   # error in atan = t * err(f)
        w = w*(1+ 3*e/e.max())
        w /= w.mean()
    return c, e.max()
for n in (5,6,7,8):
    c,e = fit(n)
    print(n, e)
c,e = fit(7)
print([f"{x:.9e}" for x in c])
np.save('/tmp/atan_c.npy', c)
import numpy as np
c = np.load('/tmp/atan_c.npy').astype(np.float32)
f32=np.float32
def fma(a,b,cc): return (a.astype(np.float64)*b.astype(np.float64)+cc.astype(np.float64)).astype(f32)
def fast_atan2(y,x, rng):
    ax,ay=np.abs(x),np.abs(y)
    mx,mn=np.maximum(ax,ay),np.minimum(ax,ay)
    t=(mn.astype(np.float64)/mx.astype(np.float64)).astype(f32)
    # perturb by up to 2 ulp like __fdividef
    t = (t * (f32(1)+ (rng.integers(-2,3,size=t.shape)).astype(f32)*f32(2**-23))).astype(f32)
    s=(t*t).astype(f32)
    p=np.full_like(s,c[7])
    for k in range(6,-1,-1): p=fma(p,s,np.full_like(s,c[k]))
    a=(t*p).astype(f32)
    a=np.where(ay>ax,(f32(np.pi/2)-a).astype(f32),a)
    a=np.where(x<0,(f32(np.pi)-a).astype(f32),a)
    return np.copysign(a,y)
rng=np.random.default_rng(0)
N=4_000_000
worst=0
for trial in range(6):
    if trial<3:
        r=10**rng.uniform(-2,3,N); th=rng.uniform(-np.pi,np.pi,N)
        qx,qy=r*np.cos(th),r*np.sin(th)
    elif trial==3:
        # near axes / diagonals
        th=np.concatenate([k*np.pi/4+rng.normal(0,1e-4,N//8) for k in range(-4,4)]); r=10**rng.uniform(-1,2,th.size)
        qx,qy=r*np.cos(th),r*np.sin(th)
    else:
        qx=rng.normal(0,20,N); qy=rng.normal(0,20,N)
    ex=np.arctan2(qy,qx)
    ap=fast_atan2(qy.astype(f32),qx.astype(f32),rng).astype(np.float64)
    err=np.abs(ap-ex); err=np.minimum(err,2*np.pi-err)
    print(trial, err.max(), 'rad =', err.max()/(2*np.pi/1024),'px cols', err.max()/np.radians(28/64),'px rows')
    worst=max(worst,err.max())
print('worst',worst)
