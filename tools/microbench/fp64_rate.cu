// B200 FP64 DADD microbenchmark: dependent-chain latency and per-SM throughput (evidence for DESIGN.md).
// build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o fp64_rate fp64_rate.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int CHAINS>
__global__ void dadd(double *out, double v, int iters) {
    double a[CHAINS];
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) a[c] = v * (c + 1);
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int c = 0; c < CHAINS; ++c) a[c] += v;
    }
    double s = 0;
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) s += a[c];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int CHAINS>
void run(const char *name, int blocks, int threads, int iters, double clock_ghz, int sms) {
    double *out;
    cudaMalloc(&out, sizeof(double) * blocks * threads);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    dadd<CHAINS><<<blocks, threads>>>(out, 1e-9, 100);
    cudaEventRecord(e0);
    dadd<CHAINS><<<blocks, threads>>>(out, 1e-9, iters);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    const double adds = (double)blocks * threads * iters * CHAINS;
    const double cyc = ms * 1e-3 * clock_ghz * 1e9;
    printf("%-34s blocks %5d x %4d thr, %d chain(s): %.3f ms, %.2f Gadd/s, %.1f thread-adds/clk/SM, %.1f clk per dependent add per warp-slot\n",
           name, blocks, threads, CHAINS, ms, adds / ms * 1e-6, adds / cyc / sms, cyc / iters / CHAINS);
    cudaFree(out);
}

int main() {
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    const double ghz = p.clockRate * 1e-6;
    const int sms = p.multiProcessorCount;
    printf("%s, %d SMs, %.3f GHz (max)\n", p.name, sms, ghz);
    run<1>("latency: 1 warp/SM, 1 chain", sms, 32, 1 << 16, ghz, sms);
    run<8>("1 warp/SM, 8 chains (ILP)", sms, 32, 1 << 14, ghz, sms);
    run<1>("4 warps/SM, 1 chain", sms, 128, 1 << 16, ghz, sms);
    run<1>("32 warps/SM, 1 chain", sms, 1024, 1 << 14, ghz, sms);
    run<4>("32 warps/SM, 4 chains", sms, 1024, 1 << 13, ghz, sms);
    run<8>("64 warps/SM, 8 chains (throughput)", 2 * sms, 1024, 1 << 12, ghz, sms);
    return 0;
}
