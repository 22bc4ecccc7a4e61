// Device -> pinned host copy rates for the result of one pipeline chunk of the host path (cpb200_forward_host):
// a pitched 2-D copy of 2507-float rows (what the tight NumPy layout needs) against contiguous copies of the same rows.
// build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o d2h_copy_bw d2h_copy_bw.cu
#include <cstdio>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); return 1; } } while (0)

int main() {
    const size_t rows = 13312 * 4, n_modes = 2507, dpitch = 2512;      // 4 chunks of the engine's 128 MB
    float *d = nullptr, *h = nullptr;
    CK(cudaMalloc(reinterpret_cast<void**>(&d), rows * dpitch * 4));
    CK(cudaMallocHost(reinterpret_cast<void**>(&h), rows * dpitch * 4));
    CK(cudaMemset(d, 1, rows * dpitch * 4));
    cudaStream_t st;
    CK(cudaStreamCreate(&st));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    const char* names[4] = {"2-D, 10028-byte rows, device pitch 10048 -> host pitch 10028 (shipped)",
                            "2-D, 10028-byte rows, device pitch 10048 -> host pitch 10048",
                            "1-D, rows x 10048 bytes (padded rows, host pitch = device pitch)",
                            "1-D, rows x 10028 bytes (tight device rows)"};
    for (int mode = 0; mode < 4; ++mode) {
        float best = 1e30f;
        for (int rep = 0; rep < 6; ++rep) {
            CK(cudaEventRecord(e0, st));
            if (mode == 0) CK(cudaMemcpy2DAsync(h, n_modes * 4, d, dpitch * 4, n_modes * 4, rows, cudaMemcpyDeviceToHost, st));
            if (mode == 1) CK(cudaMemcpy2DAsync(h, dpitch * 4, d, dpitch * 4, n_modes * 4, rows, cudaMemcpyDeviceToHost, st));
            if (mode == 2) CK(cudaMemcpyAsync(h, d, rows * dpitch * 4, cudaMemcpyDeviceToHost, st));
            if (mode == 3) CK(cudaMemcpyAsync(h, d, rows * n_modes * 4, cudaMemcpyDeviceToHost, st));
            CK(cudaEventRecord(e1, st));
            CK(cudaStreamSynchronize(st));
            float ms = 0.f;
            CK(cudaEventElapsedTime(&ms, e0, e1));
            if (rep > 0 && ms < best) best = ms;
        }
        printf("%-78s %8.3f ms  %6.2f GB/s of spectrum bytes\n", names[mode], best, rows * n_modes * 4 / best * 1e-6);
    }
    return 0;
}
