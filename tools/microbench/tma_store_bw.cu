// Microbenchmark: throughput of bulk (TMA) stores shared -> global from one SM, to compare with the 32 B/clk/SM of plain
// streaming stores (store_bw.cu): would staging the epilogue's output through shared memory lift its floor?
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tma_store_bw tma_store_bw.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__global__ void k(uint8_t* out, long long bytes_per_block, int chunk, int depth, long long* cyc) {
    extern __shared__ __align__(128) uint8_t smem[];
    for (int i = threadIdx.x; i < chunk * depth / 4; i += blockDim.x) reinterpret_cast<float*>(smem)[i] = i;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    const long long t0 = clock64();
    if (threadIdx.x == 0) {
        uint8_t* dst = out + blockIdx.x * bytes_per_block;
        const long long n = bytes_per_block / chunk;
        for (long long i = 0; i < n; ++i) {
            const uint32_t src = static_cast<uint32_t>(__cvta_generic_to_shared(smem + (i % depth) * chunk));
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" :: "l"(dst + i * chunk), "r"(src), "r"(chunk) : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            asm volatile("cp.async.bulk.wait_group.read %0;" :: "n"(3) : "memory");     // at most 4 groups reading shared memory
        }
        asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
    __syncthreads();
    const long long t1 = clock64();
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
int main() {
    long long* cyc; cudaMalloc(&cyc, 148 * sizeof(long long));
    const int depth = 4;
    for (int chunk : {4096, 16384}) for (int active : {148, 74, 1}) {
        const long long bpb = 64ll << 20;
        uint8_t* out; cudaMalloc(&out, bpb * active);
        cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, chunk * depth);
        k<<<active, 128, chunk * depth>>>(out, bpb, chunk, depth, cyc);
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        cudaEventRecord(e0); k<<<active, 128, chunk * depth>>>(out, bpb, chunk, depth, cyc); cudaEventRecord(e1);
        cudaDeviceSynchronize();
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        long long h[148]; cudaMemcpy(h, cyc, active * sizeof(long long), cudaMemcpyDeviceToHost);
        double c = 0; for (int i = 0; i < active; ++i) c += h[i]; c /= active;
        printf("%3d SMs, %5d-byte bulk stores: %.1f B/clk/SM, %.2f TB/s in total (%.3f ms) %s\n", active, chunk, bpb / c, bpb * 1.0 * active / ms / 1e9, ms, cudaGetErrorString(cudaGetLastError()));
        cudaFree(out);
    }
    return 0;
}
