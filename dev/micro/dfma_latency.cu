// Developer microbenchmark: latency and per-SM throughput of dependent / independent FP64 FMA chains on this GPU.
#include <cstdio>
#include <cuda_runtime.h>
template <int ILP>
__global__ void chain(double* out, long long* cycles, int iters, double a, double b) {
    double x[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) x[i] = threadIdx.x + i;
    long long t0 = clock64();
    for (int k = 0; k < iters; ++k) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) x[i] = fma(x[i], a, b);
    }
    long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = t1 - t0;
}
template <int ILP>
void run(int warps) {
    double* out; long long* cyc; cudaMalloc(&out, 148 * 1024 * 8); cudaMalloc(&cyc, 8);
    const int iters = 4096;
    chain<ILP><<<148, warps * 32>>>(out, cyc, iters, 1.0000001, 1e-9);
    cudaDeviceSynchronize();
    chain<ILP><<<148, warps * 32>>>(out, cyc, iters, 1.0000001, 1e-9);
    cudaDeviceSynchronize();
    long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    double per = (double)c / iters;
    printf("ILP %d warps/SM %2d: %.2f cycles per round of %d DFMA per warp -> %.2f warp-DFMA per cycle per SM (%.1f TFLOP/s at 1.965 GHz)\n", ILP, warps, per, ILP,
           ILP * warps / per, ILP * warps / per * 64 * 148 * 1.965e9 / 1e12);
    cudaFree(out); cudaFree(cyc);
}
int main() {
    for (int w : {1, 4, 8, 16, 32}) run<1>(w);
    for (int w : {1, 4, 8, 16, 32}) run<4>(w);
    for (int w : {1, 4, 16}) run<8>(w);
    return 0;
}
