"""numpy model of K1's middle pass (csrc/k1_locate.cu, row_span + scipy's acceptance rule): for
random triangles -- general, slivers up to 120 degrees long, vertices exactly on mesh nodes,
horizontal edges exactly on a mesh row, float32 coordinates -- every mesh node of the bounding
box that passes the inside test must lie within the column span the kernel tests on its row.

    python tools/locate_span_model.py SEED SECONDS

Round 2: 6 seeds x 240 s, 785,000 triangles, no node outside its span.
"""
import numpy as np, sys, time
rng=np.random.default_rng(int(sys.argv[1]))
EPS=100*2.220446049250313e-16
def tri_geom(x0,y0,x1,y1,x2,y2):
    a,b,c,d=x0-x2,x1-x2,y0-y2,y1-y2
    det=a*d-b*c
    if det==0 or det!=det: return None
    r=1.0/det
    return (x2,y2,d*r,-b*r,-c*r,a*r)
def inside(g,qx,qy):
    x2,y2,t00,t01,t10,t11=g
    d0=qx-x2; d1=qy-y2
    c0=t00*d0+t01*d1; c1=t10*d0+t11*d1; c2=1.0-c0-c1
    return (c0>=-EPS)&(c0<=1+EPS)&(c1>=-EPS)&(c1<=1+EPS)&(c2>=-EPS)&(c2<=1+EPS)
def row_span(vx,vy,y,x0,sx,i_lo,i_hi):
    tol=1e-9*(abs(y)+1.0)
    xl=np.inf; xr=-np.inf
    for e in range(3):
        ax,ay,bx,by=vx[e],vy[e],vx[(e+1)%3],vy[(e+1)%3]
        if abs(y-ay)<=tol: xl=min(xl,ax); xr=max(xr,ax)
        lo,hi=min(ay,by),max(ay,by)
        if y<lo-tol or y>hi+tol or hi==lo: continue
        x=ax+(y-ay)/(by-ay)*(bx-ax)
        x=min(max(x,min(ax,bx)),max(ax,bx))
        xl=min(xl,x); xr=max(xr,x)
    if not (xl<=xr): return None
    fa=np.floor((xl-x0)/sx)-1.0; fb=np.ceil((xr-x0)/sx)+1.0
    ia=i_lo if fa<i_lo else int(fa); ib=i_hi if fb>i_hi else int(fb)
    return (ia,ib) if ia<=ib else None
W,H=1441,721
xs=np.linspace(-180,180,W); ys=np.linspace(-90,90,H); x0=xs[0]; y0=ys[0]; sx=(xs[-1]-x0)/(W-1); sy=(ys[-1]-y0)/(H-1)
t0=time.time(); n=0; miss=0
while time.time()-t0<float(sys.argv[2]):
    k=rng.integers(0,5)
    c=np.array([rng.uniform(-170,170),rng.uniform(-85,85)])
    if k==0: v=c+rng.normal(0,rng.choice([0.2,1,5,30]),(3,2))
    elif k==1: # sliver
        d=rng.normal(0,1,2); d/=np.linalg.norm(d); L=rng.uniform(5,120)
        v=np.array([c,c+d*L,c+d*L*rng.uniform(0.2,0.8)+np.array([-d[1],d[0]])*rng.uniform(0.01,0.6)])
    elif k==2: # vertices exactly on mesh nodes / rows
        ii=rng.integers(0,W,3); jj=rng.integers(max(0,int((c[1]+90)/sy)-20),min(H,int((c[1]+90)/sy)+20),3)
        v=np.column_stack((xs[ii],ys[jj]))
    elif k==3: # horizontal edge on a mesh row
        j=rng.integers(5,H-5); v=np.array([[c[0],ys[j]],[c[0]+rng.uniform(0.3,40),ys[j]],[c[0]+rng.uniform(-10,30),ys[j]+rng.uniform(-8,8)]])
    else: # float32 coordinates
        v=(c+rng.normal(0,2,(3,2))).astype(np.float32).astype(np.float64)
    v[:,0]=np.clip(v[:,0],-180,180); v[:,1]=np.clip(v[:,1],-90,90)
    g=tri_geom(v[0,0],v[0,1],v[1,0],v[1,1],v[2,0],v[2,1])
    if g is None: continue
    xmin,xmax,ymin,ymax=v[:,0].min(),v[:,0].max(),v[:,1].min(),v[:,1].max()
    i0=int(max(0,np.floor((xmin-x0)/sx)-1)); i1=int(min(W-1,np.ceil((xmax-x0)/sx)+1))
    j0=int(max(0,np.floor((ymin-y0)/sy)-1)); j1=int(min(H-1,np.ceil((ymax-y0)/sy)+1))
    if (i1-i0+1)*(j1-j0+1)>400000: continue
    for j in range(j0,j1+1):
        ins=inside(g,xs[i0:i1+1],ys[j])
        idx=np.flatnonzero(ins)+i0
        sp=row_span(v[:,0],v[:,1],ys[j],x0,sx,i0,i1)
        if idx.size:
            if sp is None or idx.min()<sp[0] or idx.max()>sp[1]:
                miss+=1; print('MISS kind',k,'row',j,'nodes',idx.min(),idx.max(),'span',sp,v.tolist()); break
    n+=1
print('triangles',n,'misses',miss)
