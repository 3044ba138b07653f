// FP64 FMA issue rate of the CUDA cores (not the tensor pipe): nvcc -arch=sm_100a -O3 -o fp64_rate fp64_rate.cu && ./fp64_rate
#include <cstdio>
#include <cuda_runtime.h>
template <int ILP>
__global__ void k(double* out, int iters, double a, double b) {
    double x[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) x[i] = threadIdx.x + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) x[i] = fma(x[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int ILP>
void run(int threads, int blocks_per_sm) {
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    double* d; cudaMalloc(&d, size_t(sms) * blocks_per_sm * threads * 8);
    const int iters = 20000;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<ILP><<<sms * blocks_per_sm, threads>>>(d, 100, 1.0000001, 1e-9);
    cudaEventRecord(e0);
    k<ILP><<<sms * blocks_per_sm, threads>>>(d, iters, 1.0000001, 1e-9);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double fmas = double(sms) * blocks_per_sm * threads * double(iters) * ILP;
    printf("ILP %d, %d threads x %d blocks/SM: %.2f TFLOP/s FP64 (FMA = 2), %.1f FMA/clk/SM at 1.9 GHz\n", ILP, threads, blocks_per_sm,
           2 * fmas / ms / 1e9, fmas / (ms * 1e-3) / sms / 1.9e9);
    cudaFree(d);
}
int main() {
    run<1>(256, 1); run<4>(256, 1); run<8>(256, 1); run<8>(512, 1); run<8>(1024, 1); run<8>(1024, 2);
    return 0;
}
