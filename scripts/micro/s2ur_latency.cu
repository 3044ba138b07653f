// latency of special-register reads and shared-window address materialisation on B200
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(long long* out, int iters) {
    extern __shared__ double sm[];
    unsigned acc = 0;
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
        unsigned r;
        asm volatile("mov.u32 %0, %%cluster_ctaid.x;" : "=r"(r));
        acc += r + i;
    }
    long long t1 = clock64();
    unsigned acc2 = 0;
    for (int i = 0; i < iters; ++i) {
        unsigned r;
        asm volatile("mov.u32 %0, %%laneid;" : "=r"(r));
        acc2 += r + i;
    }
    long long t2 = clock64();
    // dependent LDS chain
    sm[threadIdx.x] = threadIdx.x;
    __syncthreads();
    int idx = threadIdx.x;
    long long t3 = clock64();
    for (int i = 0; i < iters; ++i) idx = int(sm[idx & 31]);
    long long t4 = clock64();
    // barrier
    for (int i = 0; i < iters; ++i) __syncthreads();
    long long t5 = clock64();
    if (threadIdx.x == 0) { out[0] = (t1 - t0) / iters; out[1] = (t2 - t1) / iters; out[2] = (t4 - t3) / iters; out[3] = (t5 - t4) / iters; out[4] = acc + acc2 + idx; }
}
int main() {
    long long* d; cudaMalloc(&d, 64);
    for (int threads : {32, 128}) {
        k<<<1, threads, 1024>>>(d, 2000);
        cudaDeviceSynchronize();
        long long h[5]; cudaMemcpy(h, d, 40, cudaMemcpyDeviceToHost);
        printf("threads %d: cluster_ctaid read %lld cyc, laneid read %lld cyc, dependent LDS(+cvt) %lld cyc, __syncthreads %lld cyc\n", threads, h[0], h[1], h[2], h[3]);
    }
    return 0;
}
