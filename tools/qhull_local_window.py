"""Can Qhull on a WINDOW around a near tie stand in for Qhull on the whole granule?  A near
co-circular quadruple is planted in a full OMI-shaped granule at a margin inside Qhull's merge
zone (3e-17 <= r <= 5e-15, r as in tools/qhull_margin.py); the diagonal Qhull picks for that
quadrilateral on the whole granule (98,640 points) is compared with the one it picks on a window
of 40 lines x 10 pixels around it (+ the granule's bounding box, so that scaling and round-off
bounds are the same), and with the exact answer.

    python tools/qhull_local_window.py SEED TRIALS

Round 2, 6 seeds x 40 trials: window == whole granule in 132 of 240 cases, whole granule == exact
in 120 of 240: both are coin flips.  Inside the merge zone Qhull's diagonal is the fan from the
merged facet's vertex of highest id, i.e. from the point its furthest-point rule added last --
a property of the whole point set that no local computation reproduces.  The fallback for a near
tie with a kept mesh node therefore stays a Qhull run on the whole granule.
"""
import sys, numpy as np, time
from fractions import Fraction as F
from scipy.spatial import Delaunay
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import synth
from oisatgmi_b200 import plan
rng=np.random.default_rng(int(sys.argv[1]))
def incircle_exact(a,b,c,d):
    ax,ay=F(a[0])-F(d[0]),F(a[1])-F(d[1]); bx,by=F(b[0])-F(d[0]),F(b[1])-F(d[1]); cx,cy=F(c[0])-F(d[0]),F(c[1])-F(d[1])
    return (ax*ax+ay*ay)*(bx*cy-cx*by)+(bx*bx+by*by)*(cx*ay-ax*cy)+(cx*cx+cy*cy)*(ax*by-bx*ay)
def orient(a,b,c):
    return (F(b[0])-F(a[0]))*(F(c[1])-F(a[1]))-(F(b[1])-F(a[1]))*(F(c[0])-F(a[0]))
def diag(simplices, quad):
    # which diagonal of the quad (ia,ib,ic,idd): edge ia-ib present (triangles (ia,ib,ic),(ia,ib,idd)) or ic-idd
    ia,ib,ic,idd=quad
    S={tuple(sorted(t)) for t in simplices.tolist() if len(set(t)&{ia,ib,ic,idd})>=3}
    if tuple(sorted((ia,ib,ic))) in S and tuple(sorted((ia,ib,idd))) in S: return 0
    if tuple(sorted((ic,idd,ia))) in S and tuple(sorted((ic,idd,ib))) in S: return 1
    return -1
N=int(sys.argv[2]); out=[]
t0=time.time()
for trial in range(N):
    node=rng.choice([0.0, 100.0, -60.0])
    lat,lon=synth.swath_geolocation(1644,60,node_lon_deg=node,rng=rng)
    lon=lon.astype(np.float64); lat=lat.astype(np.float64)
    r0=int(rng.integers(100,1500)); c0=int(rng.integers(5,45))
    m=max(np.abs(lon).max(),np.abs(lat).max())
    R,C=40,10
    sl=(slice(r0,r0+R),slice(c0,c0+C))
    idx=np.arange(lon.size).reshape(lon.shape)[sl].ravel()
    pts=np.column_stack((lon.ravel(),lat.ravel()))
    loc=pts[idx]
    tri0,ties0=plan.native_delaunay(loc[:,0],loc[:,1])
    ctr=loc.mean(0); order=np.argsort(((loc[tri0].mean(1)-ctr)**2).sum(1))
    t=tri0[order[int(rng.integers(0,6))]]
    ia,ib,ic=[int(v) for v in t]
    nb=[x for x in tri0 if ia in x and ib in x and ic not in x]
    if not nb: continue
    idd=[int(v) for v in nb[0] if v not in (ia,ib)][0]
    a,b,c=loc[ia],loc[ib],loc[ic]
    if orient(a,b,c)<0: a,b=b,a
    d=loc[idd].copy()
    target=10**rng.uniform(-16.5,-14.3); sign=rng.choice([-1,1]); want=sign*target
    A=np.array([[b[0]-a[0],b[1]-a[1]],[c[0]-a[0],c[1]-a[1]]]); rhs=0.5*np.array([b@b-a@a,c@c-a@a])
    cc=np.linalg.solve(A,rhs); rad=np.linalg.norm(a-cc); u=(d-cc)/np.linalg.norm(d-cc)
    area2=float(abs(orient(a,b,c)))
    def ratio(tt):
        p=cc+u*(rad*(1+tt)); return p, float(incircle_exact(a,b,c,p))/(m*m*area2)
    lo,hi=-1e-4,1e-4
    for _ in range(200):
        mid=0.5*(lo+hi); p,rv=ratio(mid)
        if rv>want: lo=mid
        else: hi=mid
    p,rv=ratio(0.5*(lo+hi))
    best=(abs(rv-want),p,rv)
    for dx in range(-4,5):
        for dy in range(-4,5):
            q=p.copy()
            for _ in range(abs(dx)): q[0]=np.nextafter(q[0],np.inf*np.sign(dx))
            for _ in range(abs(dy)): q[1]=np.nextafter(q[1],np.inf*np.sign(dy))
            rq=float(incircle_exact(a,b,c,q))/(m*m*area2)
            if abs(rq-want)<best[0]: best=(abs(rq-want),q,rq)
    _,p,rv=best
    if rv==0: continue
    full=pts.copy(); full[idx[idd]]=p
    ext=np.array([[lon.min(),lat.min()],[lon.min(),lat.max()],[lon.max(),lat.min()],[lon.max(),lat.max()]])
    loc2=np.vstack((full[idx],ext))
    g=Delaunay(full).simplices
    l=Delaunay(loc2).simplices
    quad_g=(idx[ia],idx[ib],idx[ic],idx[idd]); quad_l=(ia,ib,ic,idd)
    dg,dl=diag(g,quad_g),diag(l,quad_l)
    exact = 1 if rv>0 else 0      # d inside circle(a,b,c): edge ia-ib is NOT Delaunay -> diagonal ic-idd
    out.append((abs(rv),dg,dl,exact))
    print(trial,'r=%.2e'%abs(rv),'global',dg,'local',dl,'exact',exact,flush=True)
ok=[o for o in out if o[1]>=0 and o[2]>=0]
print('usable',len(ok),'global==local',sum(1 for o in ok if o[1]==o[2]),'global==exact',sum(1 for o in ok if o[1]==o[3]),'time %.0f'%(time.time()-t0))
