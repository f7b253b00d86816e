import numpy as np, sys, torch, ctypes as C
sys.path.insert(0,'/root/repo')
import astropaint_b200 as ap
ap.paint_bucket.verbose=False
from astropaint_b200 import synthetic, _engine as eng
from oracle import reference_loop as RL, profiles_np as P
L=C.CDLL('/root/repo/astropaint_b200/_native/libap_hostcheck.so')
dd=C.c_double
L.hc_qagi_b16.argtypes=[dd,dd,dd,dd,C.POINTER(dd),C.POINTER(dd),C.POINTER(C.c_int),C.POINTER(C.c_int)]
cat=ap.Catalog(synthetic.websky_like(60,seed=11,m_min=1e14)); ns=1024
dev=torch.device('cuda',0)
halos=eng.DeviceHalos(cat.data,5,dev)
n_pix,_,rmin,rmax=eng.disc_count(halos,ns,want_range=True)
nodes,n_nodes=eng.table_nodes(halos,n_pix,rmin,rmax,15,'linspace')
n_nodes[:]=15
nodes[:]=torch.linspace(0.05,3.0,15,dtype=torch.float64,device=dev)[None,:]
d=cat.data
R200,M200,z=[eng.to_device(d[k].values,dev) for k in ('R_200c','M_200c','redshift')]
vals=eng.b16_tau_nodes(halos,R200,M200,z,nodes,n_nodes).cpu().numpy()
nodes=nodes.cpu().numpy()
for h in (0,17):
    for i in range(15):
        res=dd();ae=dd();ne=C.c_int();ier=C.c_int()
        L.hc_qagi_b16(nodes[h,i],d.R_200c[h],d.M_200c[h],d.redshift[h],C.byref(res),C.byref(ae),C.byref(ne),C.byref(ier))
        o=float(P.b16_tau_2D(nodes[h,i],d.R_200c[h],d.M_200c[h],d.redshift[h]))
        print(h,i,nodes[h,i],'gpu/host-1',vals[h,i]/res.value-1,'host/oracle-1',res.value/o-1,ne.value)
