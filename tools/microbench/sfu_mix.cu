// Microbenchmark: how fast do MUFU.EX2 / MUFU.RCP and FMA-pipe instructions issue from 2 / 3 / 4 / 8 warps per SM sub-partition
// (the epilogue of the fused kernel runs 2 warps per scheduler)?  Prints results per clock per SM for a few instruction mixes.
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o sfu_mix sfu_mix.cu
#include <cstdio>
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>
__device__ __forceinline__ void split2(float a, float b, uint32_t& hi, uint32_t& lo) {
    __half2 h = __floats2half2_rn(a, b); float2 hf = __half22float2(h); __half2 l = __floats2half2_rn(a - hf.x, b - hf.y);
    hi = *reinterpret_cast<uint32_t*>(&h); lo = *reinterpret_cast<uint32_t*>(&l);
}
__device__ __forceinline__ float ex2a(float x) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float rcpa(float x) { float y; asm volatile("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
template <int MODE, int CH>
__global__ void k(float* out, int iters, long long* cyc) {
    float v[CH];
#pragma unroll
    for (int i = 0; i < CH; ++i) v[i] = 0.001f * (threadIdx.x + i);
    const float a = 0.999f + 1e-9f * blockIdx.x, b = 1e-3f;
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < CH; ++i) {
            float z = v[i];
            if (MODE == 0) z = ex2a(z);                                   // 1 MUFU
            if (MODE == 1) { z = ex2a(z); z = rcpa(z); }                  // 2 MUFU
            if (MODE == 2) { z = fmaf(z, a, b); float e = ex2a(z * a); e = rcpa(1.f + e); z = fmaf(b, e, a) * z;   // activation-like: 2 MUFU + 5 FMA-pipe
                             z = fmaf(z, a, b); z = fmaf(z, a, b); z = fmaf(z, a, b); z = fmaf(z, a, b); z = fmaf(z, a, b); z = fmaf(z, a, b); }
            if (MODE == 3) { z = fmaf(z, a, b); float e = ex2a(z * a); float d = fminf(1.f + e, 1e30f);               // 1 MUFU + reciprocal on the FMA pipe
                             float r = __int_as_float(0x7EF311C7 - __float_as_int(d)); float q = fmaf(-d, r, 1.f); r = fmaf(r, fmaf(q, q, q), r); q = fmaf(-d, r, 1.f); r = fmaf(r, q, r);
                             z = fmaf(b, r, a) * z;
                             z = fmaf(z, a, b); z = fmaf(z, a, b); z = fmaf(z, a, b); z = fmaf(z, a, b); z = fmaf(z, a, b); z = fmaf(z, a, b); }
            if (MODE == 4) { z = fmaf(z, a, b); z = fmaf(z, a, b); z = fmaf(z, a, b); z = fmaf(z, a, b); z = fmaf(z, a, b); z = fmaf(z, a, b); z = fmaf(z, a, b);
                             z = fmaf(z, a, b); z = fmaf(z, a, b); z = fmaf(z, a, b); z = fmaf(z, a, b); z = fmaf(z, a, b); z = fmaf(z, a, b); }   // 13 FMA
            if (MODE == 5) { z = fmaf(z, a, b); z = ex2a(z); }            // output-like: 1 FMA + 1 MUFU
            if (MODE == 6 || MODE == 7) { z = fmaf(z, a, b); float e = ex2a(z * a); e = rcpa(1.f + e); z = fmaf(b, e, a) * z; }   // the real activation: 2 MUFU + 5 FMA
            v[i] = z;
        }
        if (MODE == 6 || MODE == 8) {          // the real fp16 hi/lo split of pairs (F2FP, HADD2.F32 x2, FADD x2, F2FP per pair)
#pragma unroll
            for (int i = 0; i < CH; i += 2) { uint32_t hi, lo; split2(v[i], v[i + 1], hi, lo); v[i] = __uint_as_float((hi & 0x007fffffu) | 0x3f000000u); v[i + 1] = __uint_as_float((lo & 0x007fffffu) | 0x3f000000u); }
        }
    }
    const long long t1 = clock64();
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < CH; ++i) s += v[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
template <int MODE, int CH>
void run(const char* name, int threads, float* out, long long* cyc, int mufu_per_elem, int fma_per_elem) {
    const int iters = 2000, blocks = 148;
    k<MODE, CH><<<blocks, threads>>>(out, iters, cyc);
    k<MODE, CH><<<blocks, threads>>>(out, iters, cyc);
    cudaDeviceSynchronize();
    long long h[148]; cudaMemcpy(h, cyc, sizeof h, cudaMemcpyDeviceToHost);
    double c = 0; for (int i = 0; i < blocks; ++i) c += h[i]; c /= blocks;
    const double elems = double(iters) * CH * threads;
    printf("%-34s warps/SM=%2d chains=%2d: %.2f elements/clk/SM, %.2f MUFU/clk/SM, %.2f FMA-pipe thread-instr/clk/SM\n", name, threads / 32, CH,
           elems / c, elems * mufu_per_elem / c, elems * fma_per_elem / c);
}
int main() {
    float* out; long long* cyc;
    cudaMalloc(&out, 148 * 1024 * sizeof(float)); cudaMalloc(&cyc, 148 * sizeof(long long));
    for (int threads : {256, 512}) {
        run<0, 16>("ex2 only", threads, out, cyc, 1, 0);
        run<1, 16>("ex2 + rcp", threads, out, cyc, 2, 0);
        run<5, 16>("output-like (1 FMA + ex2)", threads, out, cyc, 1, 1);
        run<2, 8>("activation-like (2 MUFU + 11 FMA)", threads, out, cyc, 2, 11);
        run<2, 16>("activation-like (2 MUFU + 11 FMA)", threads, out, cyc, 2, 11);
        run<3, 16>("rcp on FMA pipe (1 MUFU + 18 FMA)", threads, out, cyc, 1, 18);
        run<4, 16>("13 FMA", threads, out, cyc, 0, 13);
        run<7, 16>("real activation, no split", threads, out, cyc, 2, 5);
        run<6, 16>("real activation + fp16 split", threads, out, cyc, 2, 10);
        run<8, 16>("fp16 split only (+2 LOP)", threads, out, cyc, 0, 5);
    }
    return 0;
}
