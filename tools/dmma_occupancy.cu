// How many warps per scheduler does the FP64 tensor pipe need?  DMMA throughput with 1 CTA per SM and 1..8 warps per SMSP,
// for the three f64 mma shapes, 8 independent accumulator sets per warp.  Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void dmma884(double &c0, double &c1, double a, double b)
{ asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b)); }
__device__ __forceinline__ void dmma1688(double *c, const double *a, const double *b)
{ asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
  : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3]) : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(b[0]), "d"(b[1])); }
__device__ __forceinline__ void dmma16816(double *c, const double *a, const double *b)
{ asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};"
  : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3]) : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]), "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3])); }
template <int SHAPE, int ILP> __global__ void k(double *out, int iters)
{
    double c[ILP][4], a[8], b[4];
    for (int i = 0; i < 8; i++) a[i] = threadIdx.x * 1e-3 + i;
    for (int i = 0; i < 4; i++) b[i] = 1.0 / (i + 1);
#pragma unroll
    for (int i = 0; i < ILP; i++) { c[i][0] = i; c[i][1] = -i; c[i][2] = 1; c[i][3] = 2; }
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < ILP; i++) {
            if (SHAPE == 0) dmma884(c[i][0], c[i][1], a[0], b[0]);
            else if (SHAPE == 1) dmma1688(c[i], a, b);
            else dmma16816(c[i], a, b);
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; i++) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int SHAPE, int ILP> void run(const char *name, double flop_per_mma, int nsm, double *out)
{
    for (int wps = 1; wps <= 8; wps *= 2) {
        const int tpb = 128 * wps, iters = 20000;
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        k<SHAPE, ILP><<<nsm, tpb>>>(out, 100); cudaDeviceSynchronize();
        cudaEventRecord(e0); k<SHAPE, ILP><<<nsm, tpb>>>(out, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        printf("%s ILP%d  %d warp(s)/SMSP: %.2f TFLOP/s\n", name, ILP, wps, flop_per_mma * ILP * iters * (double)nsm * (tpb / 32) / ms * 1e-9);
    }
}
int main()
{
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    double *out; cudaMalloc(&out, sizeof(double) * p.multiProcessorCount * 1024);
    printf("device %s SMs %d\n", p.name, p.multiProcessorCount);
    run<0, 8>("m8n8k4  ", 2.0 * 8 * 8 * 4, p.multiProcessorCount, out);
    run<0, 2>("m8n8k4  ", 2.0 * 8 * 8 * 4, p.multiProcessorCount, out);
    run<1, 8>("m16n8k8 ", 2.0 * 16 * 8 * 8, p.multiProcessorCount, out);
    run<2, 8>("m16n8k16", 2.0 * 16 * 8 * 16, p.multiProcessorCount, out);
    run<2, 2>("m16n8k16", 2.0 * 16 * 8 * 16, p.multiProcessorCount, out);
    return 0;
}
