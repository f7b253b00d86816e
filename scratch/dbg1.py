import numpy as np, sys, torch
sys.path.insert(0,'/root/repo')
import astropaint_b200 as ap
ap.paint_bucket.verbose=False
from astropaint_b200 import synthetic, _engine as eng
from oracle import reference_loop as RL, profiles_np as P
from scipy.interpolate import InterpolatedUnivariateSpline
cat=ap.Catalog(synthetic.websky_like(60,seed=11,m_min=1e14)); ns=1024
dev=torch.device('cuda',0)
halos=eng.DeviceHalos(cat.data,5,dev)
n_pix,_,rmin,rmax=eng.disc_count(halos,ns,want_range=True)
nodes,n_nodes=eng.table_nodes(halos,n_pix,rmin,rmax,15,'linspace')
d=cat.data
R200,M200,z=[eng.to_device(d[k].values,dev) for k in ('R_200c','M_200c','redshift')]
vals=eng.b16_tau_nodes(halos,R200,M200,z,nodes,n_nodes)
coef=eng.spline_build(nodes,vals,n_nodes)
discs=RL.Discs(d,ns,5)
nodes=nodes.cpu().numpy(); vals=vals.cpu().numpy(); coef=coef.cpu().numpy(); npx=n_pix.cpu().numpy()
for h in range(8):
    q=discs.pixel_index(h); R=discs.cent2pix_mpc(h,q)
    if len(R)<15: print(h,'small',len(R),npx[h]); continue
    xs=np.linspace(R.min(),R.max(),15)
    v=np.array([P.b16_tau_2D(x,d.R_200c[h],d.M_200c[h],d.redshift[h]) for x in xs])
    print(h,len(R),npx[h],'nodes',np.abs(nodes[h]/xs-1).max(),'vals',np.abs(vals[h]/v-1).max())
    sp=InterpolatedUnivariateSpline(xs,v,k=3)
    # eval pp
    X=R
    i=np.clip(np.searchsorted(nodes[h],X,side='right')-1,0,13)
    t=X-nodes[h][i]; c=coef[h][i]
    pp=c[:,0]+t*(c[:,1]+t*(c[:,2]+t*c[:,3]))
    print('   spline',np.abs(pp/sp(X)-1).max())
print('---- per halo map compare')
from astropaint_b200.profiles import Battaglia16
canvas=ap.Canvas(cat,ns,R_times=5)
ap.Painter(Battaglia16.tau_2D_interp).spray(canvas)
got=canvas.pixels
for h in range(60):
    q=discs.pixel_index(h); R=discs.cent2pix_mpc(h,q)
    if len(q)==0: continue
    w=P.b16_tau_2D_interp(R,d.R_200c[h],d.M_200c[h],d.redshift[h])
    e=np.abs(got[q]/w-1).max()
    if len(R)>=15:
        xs=np.linspace(R.min(),R.max(),15)
        v=np.array([P.b16_tau_2D(x,d.R_200c[h],d.M_200c[h],d.redshift[h]) for x in xs])
        print(h,len(q),'err',e,'nodes',np.abs(nodes[h]/xs-1).max(),'vals',np.abs(vals[h]/v-1).max(), 'Rmin/R200',R.min()/d.R_200c[h])
    else:
        print(h,len(q),'err',e)
