// Microbenchmark: DFMA throughput when one operand streams from __constant__ (uniform, compile-time offsets)
// vs. from shared memory (broadcast LDS.128), working set 12.8 KB like the order-3 operator tables.
#include <cstdio>
#include <cuda_runtime.h>
__constant__ double cM[1600];
template<int NACC> __global__ void k_const(double* out, int reps, const double* in) {
  double acc[NACC], x[8];
  for (int i=0;i<8;i++) x[i]=in[threadIdx.x+i*blockDim.x];
  #pragma unroll
  for (int i=0;i<NACC;i++) acc[i]=0;
  for (int r=0;r<reps;r++) {
    #pragma unroll
    for (int k=0;k<1600;k++) acc[k%NACC]=fma(cM[k], x[(k/NACC)&7], acc[k%NACC]);
  }
  double s=0;
  #pragma unroll
  for (int i=0;i<NACC;i++) s+=acc[i];
  out[blockIdx.x*blockDim.x+threadIdx.x]=s;
}
template<int NACC> __global__ void k_smem(double* out, int reps, const double* in, const double* gM) {
  __shared__ __align__(16) double sM[1600];
  for (int i=threadIdx.x;i<1600;i+=blockDim.x) sM[i]=gM[i];
  __syncthreads();
  double acc[NACC], x[8];
  for (int i=0;i<8;i++) x[i]=in[threadIdx.x+i*blockDim.x];
  #pragma unroll
  for (int i=0;i<NACC;i++) acc[i]=0;
  for (int r=0;r<reps;r++) {
    #pragma unroll
    for (int k=0;k<1600;k+=2) {
      double2 m=*reinterpret_cast<const double2*>(&sM[k]);
      acc[k%NACC]=fma(m.x, x[(k/NACC)&7], acc[k%NACC]);
      acc[(k+1)%NACC]=fma(m.y, x[((k+1)/NACC)&7], acc[(k+1)%NACC]);
    }
  }
  double s=0;
  #pragma unroll
  for (int i=0;i<NACC;i++) s+=acc[i];
  out[blockIdx.x*blockDim.x+threadIdx.x]=s;
}
int main(){
  double h[1600]; for(int i=0;i<1600;i++) h[i]=1.0/(i+1);
  cudaMemcpyToSymbol(cM,h,sizeof(h));
  double *out,*in,*gM; cudaMalloc(&out,8*148*16*1024); cudaMalloc(&in,8*8*1024); cudaMalloc(&gM,sizeof(h));
  cudaMemset(in,0,8*8*1024); cudaMemcpy(gM,h,sizeof(h),cudaMemcpyHostToDevice);
  int reps=200;
  for (int tpb : {128,256,384}) for (int bps : {1,2,4}) {
    int blocks=148*bps; cudaEvent_t e0,e1; cudaEventCreate(&e0); cudaEventCreate(&e1); float ms;
    k_const<20><<<blocks,tpb>>>(out,reps,in); cudaDeviceSynchronize();
    cudaEventRecord(e0); k_const<20><<<blocks,tpb>>>(out,reps,in); cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms,e0,e1);
    double fl=2.0*1600*reps*(double)blocks*tpb;
    printf("const NACC20 tpb=%d blocks/SM=%d: %.3f ms %.2f TFLOP/s\n",tpb,bps,ms,fl/ms*1e-9);
    k_smem<20><<<blocks,tpb>>>(out,reps,in,gM); cudaDeviceSynchronize();
    cudaEventRecord(e0); k_smem<20><<<blocks,tpb>>>(out,reps,in,gM); cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms,e0,e1);
    printf("smem  NACC20 tpb=%d blocks/SM=%d: %.3f ms %.2f TFLOP/s\n",tpb,bps,ms,fl/ms*1e-9);
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
