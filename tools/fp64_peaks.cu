// Microbenchmark: FP64 FMA vs FP64 tensor (DMMA) throughput and their overlap on sm_100a.
// Decides whether the volume/LIFT contractions go to DMMA (north_star: "only if ncu shows compute-bound").
#include <cstdio>
#include <cuda_runtime.h>
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1;}}while(0)

template<int ILP> __global__ void k_dfma(double* out, int iters, double a, double b) {
  double acc[ILP];
  #pragma unroll
  for (int i=0;i<ILP;i++) acc[i]=threadIdx.x*1e-3+i;
  for (int it=0; it<iters; it++) {
    #pragma unroll
    for (int i=0;i<ILP;i++) acc[i]=fma(acc[i],a,b);
  }
  double s=0;
  #pragma unroll
  for (int i=0;i<ILP;i++) s+=acc[i];
  out[blockIdx.x*blockDim.x+threadIdx.x]=s;
}

__device__ __forceinline__ void dmma884(double& c0,double& c1,double a,double b){
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n":"+d"(c0),"+d"(c1):"d"(a),"d"(b));
}
__device__ __forceinline__ void dmma1688(double* c,const double* a,const double* b){
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
   :"+d"(c[0]),"+d"(c[1]),"+d"(c[2]),"+d"(c[3]):"d"(a[0]),"d"(a[1]),"d"(a[2]),"d"(a[3]),"d"(b[0]),"d"(b[1]));
}
__device__ __forceinline__ void dmma16816(double* c,const double* a,const double* b){
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};\n"
   :"+d"(c[0]),"+d"(c[1]),"+d"(c[2]),"+d"(c[3]):"d"(a[0]),"d"(a[1]),"d"(a[2]),"d"(a[3]),"d"(a[4]),"d"(a[5]),"d"(a[6]),"d"(a[7]),"d"(b[0]),"d"(b[1]),"d"(b[2]),"d"(b[3]));
}
template<int ILP> __global__ void k_dmma884(double* out,int iters){
  double c[ILP][2]; double a=threadIdx.x*1e-3, b=1.0+threadIdx.x*1e-4;
  #pragma unroll
  for(int i=0;i<ILP;i++){c[i][0]=i;c[i][1]=-i;}
  for(int it=0;it<iters;it++){
    #pragma unroll
    for(int i=0;i<ILP;i++) dmma884(c[i][0],c[i][1],a,b);
  }
  double s=0;
  #pragma unroll
  for(int i=0;i<ILP;i++) s+=c[i][0]+c[i][1];
  out[blockIdx.x*blockDim.x+threadIdx.x]=s;
}
template<int ILP> __global__ void k_dmma1688(double* out,int iters){
  double c[ILP][4]; double a[4],b[2];
  for(int i=0;i<4;i++)a[i]=threadIdx.x*1e-3+i; b[0]=1.0;b[1]=0.5;
  #pragma unroll
  for(int i=0;i<ILP;i++){c[i][0]=i;c[i][1]=-i;c[i][2]=1;c[i][3]=2;}
  for(int it=0;it<iters;it++){
    #pragma unroll
    for(int i=0;i<ILP;i++) dmma1688(c[i],a,b);
  }
  double s=0;
  #pragma unroll
  for(int i=0;i<ILP;i++) s+=c[i][0]+c[i][1]+c[i][2]+c[i][3];
  out[blockIdx.x*blockDim.x+threadIdx.x]=s;
}
template<int ILP> __global__ void k_dmma16816(double* out,int iters){
  double c[ILP][4]; double a[8],b[4];
  for(int i=0;i<8;i++)a[i]=threadIdx.x*1e-3+i; for(int i=0;i<4;i++)b[i]=1.0/(i+1);
  #pragma unroll
  for(int i=0;i<ILP;i++){c[i][0]=i;c[i][1]=-i;c[i][2]=1;c[i][3]=2;}
  for(int it=0;it<iters;it++){
    #pragma unroll
    for(int i=0;i<ILP;i++) dmma16816(c[i],a,b);
  }
  double s=0;
  #pragma unroll
  for(int i=0;i<ILP;i++) s+=c[i][0]+c[i][1]+c[i][2]+c[i][3];
  out[blockIdx.x*blockDim.x+threadIdx.x]=s;
}
// mixed: half the warps DFMA, half DMMA
__global__ void k_mixed(double* out,int iters,double fa,double fb){
  int warp=threadIdx.x>>5; double s=0;
  if(warp&1){
    double acc[8];
    #pragma unroll
    for(int i=0;i<8;i++)acc[i]=threadIdx.x*1e-3+i;
    for(int it=0;it<iters;it++){
      #pragma unroll
      for(int i=0;i<8;i++)acc[i]=fma(acc[i],fa,fb);
    }
    #pragma unroll
    for(int i=0;i<8;i++)s+=acc[i];
  } else {
    double c[8][4]; double a[4],b[2];
    for(int i=0;i<4;i++)a[i]=threadIdx.x*1e-3+i; b[0]=1.0;b[1]=0.5;
    #pragma unroll
    for(int i=0;i<8;i++){c[i][0]=i;c[i][1]=-i;c[i][2]=1;c[i][3]=2;}
    for(int it=0;it<iters;it++){
      #pragma unroll
      for(int i=0;i<8;i++) dmma1688(c[i],a,b);
    }
    #pragma unroll
    for(int i=0;i<8;i++) s+=c[i][0]+c[i][1]+c[i][2]+c[i][3];
  }
  out[blockIdx.x*blockDim.x+threadIdx.x]=s;
}
// shared-memory bandwidth: LDS.64 / LDS.128 conflict-free
template<int VEC> __global__ void k_lds(double* out,int iters){
  extern __shared__ double sm[];
  for(int i=threadIdx.x;i<4096;i+=blockDim.x) sm[i]=i;
  __syncthreads();
  double s=0; int base=(threadIdx.x*VEC)&4095;
  for(int it=0;it<iters;it++){
    #pragma unroll
    for(int u=0;u<8;u++){
      int idx=(base+u*256*VEC)&4095;
      if(VEC==1) s+=sm[idx];
      else { double2 v=*reinterpret_cast<double2*>(&sm[idx]); s+=v.x+v.y; }
    }
    base=(base+32*VEC)&4095;
  }
  out[blockIdx.x*blockDim.x+threadIdx.x]=s;
}

template<class F> float timeit(F f){
  cudaEvent_t e0,e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  f(); cudaDeviceSynchronize();
  cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms,e0,e1); return ms;
}
int main(){
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p,0));
  printf("device %s SMs %d clock %d kHz\n",p.name,p.multiProcessorCount,p.clockRate);
  int nsm=p.multiProcessorCount; double* out; CK(cudaMalloc(&out,sizeof(double)*nsm*8*1024));
  int iters=20000;
  for(int tpb: {256,512,1024}){
    int blocks=nsm*(2048/tpb);
    float ms=timeit([&]{k_dfma<8><<<blocks,tpb>>>(out,iters,1.0000001,1e-9);});
    double fl=2.0*8*iters*(double)blocks*tpb;
    printf("DFMA ILP8 tpb=%d: %.3f ms  %.2f TFLOP/s\n",tpb,ms,fl/ms*1e-9);
  }
  {int tpb=256,blocks=nsm*4; float ms=timeit([&]{k_dfma<16><<<blocks,tpb>>>(out,iters,1.0000001,1e-9);});
   printf("DFMA ILP16 tpb=256x4: %.3f ms %.2f TFLOP/s\n",ms,2.0*16*iters*(double)blocks*tpb/ms*1e-9);}
  for(int tpb: {128,256,512}){
    int blocks=nsm*(1024/tpb);
    float ms=timeit([&]{k_dmma884<8><<<blocks,tpb>>>(out,iters);});
    double fl=2.0*8*8*4*8*iters*(double)blocks*(tpb/32);
    printf("DMMA m8n8k4 ILP8 tpb=%d: %.3f ms  %.2f TFLOP/s\n",tpb,ms,fl/ms*1e-9);
    ms=timeit([&]{k_dmma1688<8><<<blocks,tpb>>>(out,iters);});
    fl=2.0*16*8*8*8*iters*(double)blocks*(tpb/32);
    printf("DMMA m16n8k8 ILP8 tpb=%d: %.3f ms  %.2f TFLOP/s\n",tpb,ms,fl/ms*1e-9);
    ms=timeit([&]{k_dmma16816<8><<<blocks,tpb>>>(out,iters);});
    fl=2.0*16*8*16*8*iters*(double)blocks*(tpb/32);
    printf("DMMA m16n8k16 ILP8 tpb=%d: %.3f ms  %.2f TFLOP/s\n",tpb,ms,fl/ms*1e-9);
  }
  {int tpb=512,blocks=nsm*2; float ms=timeit([&]{k_mixed<<<blocks,tpb>>>(out,iters,1.0000001,1e-9);});
   double fl_f=2.0*8*iters*(double)blocks*(tpb/2), fl_t=2.0*16*8*8*8*iters*(double)blocks*(tpb/64);
   printf("MIXED tpb=512: %.3f ms  DFMA %.2f + DMMA %.2f = %.2f TFLOP/s\n",ms,fl_f/ms*1e-9,fl_t/ms*1e-9,(fl_f+fl_t)/ms*1e-9);}
  {int tpb=1024,blocks=nsm*2; 
   float ms=timeit([&]{k_lds<1><<<blocks,tpb,32768>>>(out,iters);});
   printf("LDS.64 : %.3f ms  %.1f B/clk/SM-equivalent total %.2f TB/s\n",ms,0.0,8.0*8*iters*(double)blocks*tpb/ms*1e-9*1e-3);
   ms=timeit([&]{k_lds<2><<<blocks,tpb,32768>>>(out,iters);});
   printf("LDS.128: %.3f ms  total %.2f TB/s\n",ms,16.0*8*iters*(double)blocks*tpb/ms*1e-9*1e-3);}
  cudaFree(out); return 0;
}
