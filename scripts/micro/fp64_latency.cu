// FP64 vector-pipe microbenchmark for B200: dependent DFMA/DADD latency and issue throughput.
#include <cstdio>
#include <cuda_runtime.h>
template <int ILP>
__global__ void dfma_chain(double* out, int iters, double a, double b) {
    double x[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) x[i] = threadIdx.x * 1e-3 + i;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) x[i] = fma(x[i], a, b);
    }
    long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = double(t1 - t0) / iters;
}
template <int ILP>
void run(int warps) {
    double* d; cudaMalloc(&d, 1 << 20);
    dfma_chain<ILP><<<1, warps * 32>>>(d, 4096, 0.999, 1e-3);
    dfma_chain<ILP><<<1, warps * 32>>>(d, 4096, 0.999, 1e-3);
    cudaDeviceSynchronize();
    double h; cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
    printf("warps/CTA %2d ILP %d: %.1f cycles per iteration (%.2f cycles per DFMA per warp)\n", warps, ILP, h, h / ILP);
    cudaFree(d);
}
int main() {
    run<1>(1); run<2>(1); run<4>(1); run<8>(1); run<16>(1);
    run<1>(4); run<4>(4); run<8>(4); run<8>(8); run<8>(16); run<8>(32);
    return 0;
}
