// Which limit hits a 1600-DFMA unrolled body: instruction cache, constant cache, or LDS broadcast rate?
#include <cstdio>
#include <cuda_runtime.h>
__constant__ double cM[1600];
template<int NK> __global__ void k_const(double* out, int reps, const double* in) {
  double acc[20], x[8];
  for (int i=0;i<8;i++) x[i]=in[threadIdx.x+i*blockDim.x];
  #pragma unroll
  for (int i=0;i<20;i++) acc[i]=0;
  for (int r=0;r<reps;r++) {
    #pragma unroll
    for (int k=0;k<NK;k++) acc[k%20]=fma(cM[k], x[(k/20)&7], acc[k%20]);
  }
  double s=0;
  #pragma unroll
  for (int i=0;i<20;i++) s+=acc[i];
  out[blockIdx.x*blockDim.x+threadIdx.x]=s;
}
template<int NK> __global__ void k_reg(double* out, int reps, const double* in) {
  double acc[20], x[8];
  for (int i=0;i<8;i++) x[i]=in[threadIdx.x+i*blockDim.x];
  #pragma unroll
  for (int i=0;i<20;i++) acc[i]=i;
  for (int r=0;r<reps;r++) {
    #pragma unroll
    for (int k=0;k<NK;k++) acc[k%20]=fma(x[(k*7)&7], x[(k/20)&7], acc[k%20]);
  }
  double s=0;
  #pragma unroll
  for (int i=0;i<20;i++) s+=acc[i];
  out[blockIdx.x*blockDim.x+threadIdx.x]=s;
}
template<int NK> __global__ void k_smem(double* out, int reps, const double* in, const double* gM) {
  __shared__ __align__(16) double sM[1600];
  for (int i=threadIdx.x;i<1600;i+=blockDim.x) sM[i]=gM[i];
  __syncthreads();
  double acc[20], x[8];
  for (int i=0;i<8;i++) x[i]=in[threadIdx.x+i*blockDim.x];
  #pragma unroll
  for (int i=0;i<20;i++) acc[i]=0;
  for (int r=0;r<reps;r++) {
    #pragma unroll
    for (int k=0;k<NK;k+=2) {
      double2 m=*reinterpret_cast<const double2*>(&sM[k]);
      acc[k%20]=fma(m.x, x[(k/20)&7], acc[k%20]);
      acc[(k+1)%20]=fma(m.y, x[((k+1)/20)&7], acc[(k+1)%20]);
    }
  }
  double s=0;
  #pragma unroll
  for (int i=0;i<20;i++) s+=acc[i];
  out[blockIdx.x*blockDim.x+threadIdx.x]=s;
}
template<class F> void run(const char* name, int nk, int reps, int tpb, int blocks, F f){
  cudaEvent_t e0,e1; cudaEventCreate(&e0); cudaEventCreate(&e1); float ms;
  f(); cudaDeviceSynchronize();
  cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms,e0,e1);
  double fl=2.0*nk*reps*(double)blocks*tpb;
  printf("%-6s NK=%4d tpb=%d blocks=%d: %.3f ms %.2f TFLOP/s  (%s)\n",name,nk,tpb,blocks,ms,fl/ms*1e-9,cudaGetErrorString(cudaGetLastError()));
}
int main(){
  double h[1600]; for(int i=0;i<1600;i++) h[i]=1.0/(i+1);
  cudaMemcpyToSymbol(cM,h,sizeof(h));
  double *out,*in,*gM; cudaMalloc(&out,8*148*16*1024); cudaMalloc(&in,8*8*1024); cudaMalloc(&gM,sizeof(h));
  cudaMemset(in,0,8*8*1024); cudaMemcpy(gM,h,sizeof(h),cudaMemcpyHostToDevice);
  int tpb=256, blocks=148*2;
  run("const",160,2000,tpb,blocks,[&]{k_const<160><<<blocks,tpb>>>(out,2000,in);});
  run("const",400,800,tpb,blocks,[&]{k_const<400><<<blocks,tpb>>>(out,800,in);});
  run("const",800,400,tpb,blocks,[&]{k_const<800><<<blocks,tpb>>>(out,400,in);});
  run("const",1600,200,tpb,blocks,[&]{k_const<1600><<<blocks,tpb>>>(out,200,in);});
  run("reg",160,2000,tpb,blocks,[&]{k_reg<160><<<blocks,tpb>>>(out,2000,in);});
  run("reg",800,400,tpb,blocks,[&]{k_reg<800><<<blocks,tpb>>>(out,400,in);});
  run("reg",1600,200,tpb,blocks,[&]{k_reg<1600><<<blocks,tpb>>>(out,200,in);});
  run("reg",3200,100,tpb,blocks,[&]{k_reg<3200><<<blocks,tpb>>>(out,100,in);});
  run("reg",6400,50,tpb,blocks,[&]{k_reg<6400><<<blocks,tpb>>>(out,50,in);});
  run("smem",160,2000,tpb,blocks,[&]{k_smem<160><<<blocks,tpb>>>(out,2000,in,gM);});
  run("smem",800,400,tpb,blocks,[&]{k_smem<800><<<blocks,tpb>>>(out,400,in,gM);});
  run("smem",1600,200,tpb,blocks,[&]{k_smem<1600><<<blocks,tpb>>>(out,200,in,gM);});
  return 0;
}
