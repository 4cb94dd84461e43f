// DMMA (mma.sync.m8n8k4.f64) pipe characteristics on B200: throughput vs warps per SM and
// independent accumulator chains per warp.  Build: nvcc -gencode arch=compute_100a,code=sm_100a
#include <cstdio>
#include <cuda_runtime.h>
template <int ILP>
__global__ void k(double *out, int iters, double a, double b) {
    double c[ILP][2];
#pragma unroll
    for (int i = 0; i < ILP; ++i) { c[i][0] = threadIdx.x; c[i][1] = i; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                         : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += c[i][0] + c[i][1];
    if (s == 1.2345) out[0] = s;
}
template <int ILP>
void run(int warps, int sms) {
    double *o; cudaMalloc(&o, 8);
    const int iters = 20000 / ILP * 8;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e9;
    for (int r = 0; r < 3; ++r) {
        cudaEventRecord(e0);
        k<ILP><<<sms, warps * 32>>>(o, iters, 1.0000001, 1e-9);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (r) best = ms < best ? ms : best;
    }
    double fl = (double)sms * warps * ILP * iters * 512.0;
    printf("warps/SM %2d  ILP %d : %6.2f TFLOP/s  (%.1f clk per DMMA per SMSP at 1.965 GHz)\n", warps, ILP,
           fl / best / 1e9, best * 1e-3 * 1.965e9 / ((double)warps / 4 * ILP * iters));
    cudaFree(o);
}
int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int sms = p.multiProcessorCount;
    for (int w : {4, 8, 16, 32}) { run<1>(w, sms); run<2>(w, sms); run<4>(w, sms); run<8>(w, sms); }
    return 0;
}
