#define AP_QAGI_TRACE
#include <cstdio>
#include "../astropaint_b200/csrc/common.cuh"
#include "../astropaint_b200/csrc/profiles.cuh"
using namespace ap;
__global__ void k(double R){ double ctx[kCtx]; b16_setup(1.3,4e14,0.7,ctx); QagiResult q; double v=b16_tau_2D(ctx,R,&q); printf("DEV %.17g %.17g %d\n",v,q.abserr,q.neval);} 
int main(){ double R=0.05; double ctx[kCtx]; b16_setup(1.3,4e14,0.7,ctx); QagiResult q; printf("HOST\n"); double v=b16_tau_2D(ctx,R,&q); printf("HOST %.17g %.17g %d\n",v,q.abserr,q.neval);
 printf("DEVICE\n"); k<<<1,1>>>(R); cudaDeviceSynchronize(); return 0;}
