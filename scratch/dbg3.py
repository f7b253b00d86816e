import numpy as np, sys, torch, ctypes as C
sys.path.insert(0,'/root/repo')
import astropaint_b200 as ap
from astropaint_b200 import _capi
from astropaint_b200._capi import ptr
L2=_capi.lib()
L=C.CDLL('/root/repo/astropaint_b200/_native/libap_hostcheck.so')
dd=C.c_double
L.hc_qagi_b16.argtypes=[dd,dd,dd,dd,C.POINTER(dd),C.POINTER(dd),C.POINTER(C.c_int),C.POINTER(C.c_int)]
dev=torch.device('cuda',0)
n=15
R=torch.linspace(0.05,3.0,n,dtype=torch.float64,device=dev)
R200=torch.full((n,),1.3,dtype=torch.float64,device=dev); M=torch.full((n,),4e14,dtype=torch.float64,device=dev); z=torch.full((n,),0.7,dtype=torch.float64,device=dev)
res=torch.empty(n,dtype=torch.float64,device=dev); ae=torch.empty_like(res); ne=torch.empty(n,dtype=torch.int32,device=dev)
_capi.check(L2.ap_b16_tau_eval(n,ptr(R),ptr(R200),ptr(M),ptr(z),ptr(res),ptr(ae),ptr(ne),None))
torch.cuda.synchronize()
for i in range(n):
    r=dd();a=dd();k=C.c_int();ier=C.c_int()
    L.hc_qagi_b16(float(R[i]),1.3,4e14,0.7,C.byref(r),C.byref(a),C.byref(k),C.byref(ier))
    print(i,float(R[i]),float(res[i])/r.value-1,float(ae[i])/a.value-1,int(ne[i]),k.value)
