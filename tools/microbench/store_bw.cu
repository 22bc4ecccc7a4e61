// Microbenchmark: streaming-store throughput of one SM (st.global.cs.v8.f32, full 128-byte lines, 8 / 16 warps per SM,
// every SM busy) - the floor of the fused kernel's output epilogue, which writes 128 KB per accumulator chunk.
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o store_bw store_bw.cu
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(float* out, long long floats_per_block, int iters, long long* cyc) {
    float* base = out + blockIdx.x * floats_per_block;
    const float y[8] = {1, 2, 3, 4, 5, 6, 7, 8};
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        float* p = base + (static_cast<long long>(it) * blockDim.x + threadIdx.x) * 8;
        asm volatile("st.global.cs.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
                     :: "l"(p), "f"(y[0]), "f"(y[1]), "f"(y[2]), "f"(y[3]), "f"(y[4]), "f"(y[5]), "f"(y[6]), "f"(y[7]) : "memory");
    }
    __syncthreads();
    const long long t1 = clock64();
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
int main() {
    const int blocks = 148;
    long long* cyc; cudaMalloc(&cyc, blocks * sizeof(long long));
    for (int threads : {256, 512}) for (int active : {148, 74, 16, 1}) {
        const int iters = 4096;
        const long long fpb = static_cast<long long>(iters) * threads * 8;
        float* out; cudaMalloc(&out, fpb * active * sizeof(float));
        k<<<active, threads>>>(out, fpb, iters, cyc);
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        cudaEventRecord(e0); k<<<active, threads>>>(out, fpb, iters, cyc); cudaEventRecord(e1);
        cudaDeviceSynchronize();
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        long long h[148]; cudaMemcpy(h, cyc, active * sizeof(long long), cudaMemcpyDeviceToHost);
        double c = 0; for (int i = 0; i < active; ++i) c += h[i]; c /= active;
        printf("%3d SMs x %2d warps: %.1f B/clk/SM, %.2f TB/s in total (%.3f ms)\n", active, threads / 32, fpb * 4.0 / c, fpb * 4.0 * active / ms / 1e9, ms);
        cudaFree(out);
    }
    return 0;
}
