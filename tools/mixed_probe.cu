// Do the FP64 FMA pipe and the FP64 tensor (DMMA) pipe of B200 run concurrently?
// Each warp issues NF independent DFMA chains and NM independent DMMA chains per iteration.
#include <cstdio>
#include <cuda_runtime.h>
template <int NF, int NM>
__global__ void k(double *out, int iters, double a, double b) {
    double f[NF > 0 ? NF : 1], c[NM > 0 ? NM : 1][2];
#pragma unroll
    for (int i = 0; i < NF; ++i) f[i] = threadIdx.x + i;
#pragma unroll
    for (int i = 0; i < NM; ++i) { c[i][0] = threadIdx.x; c[i][1] = i; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < (NF > NM ? NF : NM); ++i) {
            if (i < NM)
                asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                             : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
            if (i < NF) f[i] = fma(f[i], a, b);
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < NF; ++i) s += f[i];
#pragma unroll
    for (int i = 0; i < NM; ++i) s += c[i][0] + c[i][1];
    if (s == 1.2345) out[0] = s;
}
template <int NF, int NM>
void run(int warps, int sms) {
    double *o; cudaMalloc(&o, 8);
    const int iters = 20000;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e9;
    for (int r = 0; r < 3; ++r) {
        cudaEventRecord(e0);
        k<NF, NM><<<sms, warps * 32>>>(o, iters, 1.0000001, 1e-9);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (r) best = ms < best ? ms : best;
    }
    double fl_f = (double)sms * warps * NF * iters * 64.0, fl_m = (double)sms * warps * NM * iters * 512.0;
    printf("warps/SM %2d  DFMA chains %2d  DMMA chains %d : DFMA %6.2f + DMMA %6.2f = %6.2f TFLOP/s\n", warps, NF, NM,
           fl_f / best / 1e9, fl_m / best / 1e9, (fl_f + fl_m) / best / 1e9);
    cudaFree(o);
}
int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int sms = p.multiProcessorCount;
    for (int w : {8, 16}) {
        run<8, 0>(w, sms); run<0, 2>(w, sms); run<8, 1>(w, sms); run<8, 2>(w, sms); run<16, 2>(w, sms); run<4, 2>(w, sms);
    }
    return 0;
}
