// Does the FP64 pipe pay for switching between DMMA and DFMA?  One CTA per SM, W warps per scheduler; every warp runs
// REP x { ND independent DMMA.8x8x4 ; NF independent DFMA } and the elapsed cycles per iteration are compared with the
// sum of the two pure streams (16 pipe cycles per DMMA, 2 per DFMA warp-instruction, per scheduler).
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/dmma_dfma_mix tools/dmma_dfma_mix.cu
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void dmma884(double &c0, double &c1, double a, double b)
{ asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b)); }
__device__ __forceinline__ void dfma(double &c, double a, double b)
{ asm volatile("fma.rn.f64 %0, %1, %2, %0;" : "+d"(c) : "d"(a), "d"(b)); }

template <int ND, int NF> __global__ void k(double *out, int iters, long long *cyc)
{
    double c[18][2], f[14];
    const double a = threadIdx.x * 1e-3, b = 1.0 + threadIdx.x * 1e-4;
#pragma unroll
    for (int i = 0; i < 18; i++) { c[i][0] = i; c[i][1] = -i; }
#pragma unroll
    for (int i = 0; i < 14; i++) f[i] = i;
    const long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < ND; i++) dmma884(c[i % 18][0], c[i % 18][1], a, b);
#pragma unroll
        for (int i = 0; i < NF; i++) dfma(f[i % 14], a, b);
    }
    const long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int i = 0; i < 18; i++) s += c[i][0] + c[i][1];
#pragma unroll
    for (int i = 0; i < 14; i++) s += f[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
template <int ND, int NF> void run(int nsm, double *out, long long *cyc)
{
    for (int w = 1; w <= 3; w++) {
        const int iters = 4000;
        k<ND, NF><<<nsm, 128 * w>>>(out, 10, cyc); cudaDeviceSynchronize();
        k<ND, NF><<<nsm, 128 * w>>>(out, iters, cyc); cudaDeviceSynchronize();
        long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
        const double per = (double)h / iters, ideal = w * (16.0 * ND + 2.0 * NF);
        printf("DMMA x%3d + DFMA x%3d per iteration, %d warp(s)/scheduler: %8.1f cycles, pure streams would need %7.1f  (x%.2f)\n", ND, NF, w, per, ideal, per / ideal);
    }
}
int main()
{
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    double *out; long long *cyc; cudaMalloc(&out, sizeof(double) * p.multiProcessorCount * 1024); cudaMalloc(&cyc, 8);
    printf("device %s\n", p.name);
    run<18, 0>(p.multiProcessorCount, out, cyc);
    run<0, 42>(p.multiProcessorCount, out, cyc);
    run<18, 42>(p.multiProcessorCount, out, cyc);
    run<36, 84>(p.multiProcessorCount, out, cyc);
    run<72, 168>(p.multiProcessorCount, out, cyc);
    run<6, 14>(p.multiProcessorCount, out, cyc);
    run<1, 2>(p.multiProcessorCount, out, cyc);
    return 0;
}
