// CUDA-core fp32 path (CPB200_PREC_FP32_SIMT): one dense layer per launch with the layer's
// epilogue fused.  It exists as the on-GPU fp32 truth for the tensor-core path and as the
// selectable exact-fp32 numeric format; it is not the fast path.
//
// Math follows reference cosmopower/cosmopower_NN.py:371-384 and
// cosmopower/cosmopower_PCAplusNN.py:368-381.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace cpb {

enum Epilogue : int {
    EPI_ACT = 0,     // h = (beta + (1-beta) * sigmoid(alpha * a)) * a,  a = acc + b   (NN.py:375-378)
    EPI_AFFINE = 1,  // h = acc * scale + shift   (PCA coefficient rescale, PCAplusNN.py:381, inner part)
    EPI_OUT = 2,     // y = acc * scale + shift, optional 10**y   (NN.py:381-384,425)
    EPI_QUADFORM = 3 // tensor-core path only: row-wise dot of the accumulator with the operand rows (Planck-lite chi^2)
};

constexpr int SIMT_BM = 64, SIMT_BN = 64, SIMT_BK = 16;

// C[M x N] = epi(A'[M x K] @ W[K x N]);  A' = (A - mean) / std when in_mean != nullptr
// (parameter standardisation, NN.py:371), else A.
template <int EPI>
__global__ void __launch_bounds__(256)
dense_simt_kernel(const float* __restrict__ A, int64_t lda, const float* __restrict__ W,
                  int K, int N, float* __restrict__ C, int64_t ldc, int64_t M,
                  const float* __restrict__ p0, const float* __restrict__ p1,
                  const float* __restrict__ p2, int ten_to, int accumulate,
                  const float* __restrict__ in_mean, const float* __restrict__ in_std)
{
    __shared__ float As[SIMT_BK][SIMT_BM + 4];
    __shared__ float Ws[SIMT_BK][SIMT_BN + 4];
    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;
    const int64_t row0 = (int64_t)blockIdx.y * SIMT_BM;
    const int col0 = blockIdx.x * SIMT_BN;

    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    for (int k0 = 0; k0 < K; k0 += SIMT_BK) {
        {   // A tile: 64 rows x 16 k, 4 elements per thread
            const int r = tid >> 2, kq = (tid & 3) * 4;
            const int64_t gr = row0 + r;
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int gk = k0 + kq + u;
                float v = 0.f;
                if (gr < M && gk < K) {
                    v = A[gr * lda + gk];
                    if (in_mean != nullptr) v = (v - in_mean[gk]) / in_std[gk];
                }
                As[kq + u][r] = v;
            }
        }
        {   // W tile: 16 k x 64 n
            const int kk = tid >> 4, nq = (tid & 15) * 4;
            const int gk = k0 + kk;
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int gn = col0 + nq + u;
                Ws[kk][nq + u] = (gk < K && gn < N) ? W[(int64_t)gk * N + gn] : 0.f;
            }
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < SIMT_BK; ++kk) {
            float a[4], w[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) a[i] = As[kk][ty * 4 + i];
#pragma unroll
            for (int j = 0; j < 4; ++j) w[j] = Ws[kk][tx * 4 + j];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], w[j], acc[i][j]);
        }
        __syncthreads();
    }

#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int64_t gr = row0 + ty * 4 + i;
        if (gr >= M) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int gn = col0 + tx * 4 + j;
            if (gn >= N) continue;
            float v = acc[i][j];
            if (EPI == EPI_ACT) {
                const float a = v + p0[gn];
                const float alpha = p1[gn], beta = p2[gn];
                const float s = 1.f / (1.f + expf(-alpha * a));      // exp(+big)=inf -> s=0, as numpy
                v = (beta + (1.f - beta) * s) * a;
            } else {
                v = fmaf(v, p0[gn], p1[gn]);
                if (EPI == EPI_OUT && accumulate) v += C[gr * ldc + gn];      // fused sum of log-spectra
                if (EPI == EPI_OUT && ten_to) v = exp10f(v);
            }
            C[gr * ldc + gn] = v;
        }
    }
}

// Tiny batches (a sampler's single point, up to SMALL_M rows): the layer is a handful of dot products per output column, so
// the tile kernel above - 8 blocks for a 512-wide layer, each walking K in 16-deep steps of dependent global loads - spends
// 30+ us per layer waiting.  Here a warp owns one output column: it streams that column's K weights once (TRANSPOSED copy
// Wt[N][Kp], Kp = K padded to a multiple of 4 with zeros: 16-byte coalesced loads), every lane accumulates its share for
// all rows from the activations in shared memory, five shuffles per row finish the sums, and lane r applies the epilogue
// of row r.  N / 4 blocks per layer: the whole device reads the layer's weights at once (L2-resident after the first call).
constexpr int SMALL_M = 8, SMALL_KP_MAX = 1024, SMALL_WARPS = 4;

template <int EPI>
__global__ void __launch_bounds__(32 * SMALL_WARPS)
dense_small_kernel(const float* __restrict__ A, int64_t lda, const float* __restrict__ Wt, int K, int Kp, int N,
                   float* __restrict__ C, int64_t ldc, int M,
                   const float* __restrict__ p0, const float* __restrict__ p1, const float* __restrict__ p2,
                   int ten_to, int accumulate, const float* __restrict__ in_mean, const float* __restrict__ in_std)
{
    __shared__ __align__(16) float As[SMALL_M][SMALL_KP_MAX];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    // Programmatic dependent launch: the layers of a forward pass are launched back to back with the "programmatic stream
    // serialization" attribute, so this grid may start while the previous layer's is still draining.  Everything that does not
    // depend on that layer's output - this warp's weight column, the largest read of the kernel - is fetched first; then
    // griddepcontrol.wait blocks until the previous grid has completed and its writes are visible.  (Without the attribute
    // both instructions are no-ops.)
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    const int j = blockIdx.x * SMALL_WARPS + warp;
    constexpr int WREG = SMALL_KP_MAX / 4 / 32;
    float4 wreg[WREG];
    {
        const float4* w4 = reinterpret_cast<const float4*>(Wt + static_cast<int64_t>(j < N ? j : 0) * Kp);
#pragma unroll
        for (int u = 0; u < WREG; ++u) {
            const int k4 = lane + 32 * u;
            wreg[u] = k4 < Kp / 4 ? w4[k4] : make_float4(0.f, 0.f, 0.f, 0.f);
        }
    }
    asm volatile("griddepcontrol.wait;" ::: "memory");
    for (int i = tid; i < M * Kp; i += 32 * SMALL_WARPS) {
        const int r = i / Kp, k = i - r * Kp;
        float v = 0.f;
        if (k < K) {
            v = A[r * lda + k];
            if (in_mean != nullptr) v = (v - in_mean[k]) / in_std[k];
        }
        As[r][k] = v;
    }
    __syncthreads();
    if (j >= N) return;
    float acc[SMALL_M];
#pragma unroll
    for (int r = 0; r < SMALL_M; ++r) acc[r] = 0.f;
#pragma unroll
    for (int u = 0; u < WREG; ++u) {
        const int k4 = lane + 32 * u;
        if (k4 < Kp / 4) {
            const float4 w = wreg[u];
#pragma unroll
            for (int r = 0; r < SMALL_M; ++r) {
                if (r < M) {
                    const float4 a = *reinterpret_cast<const float4*>(&As[r][4 * k4]);
                    acc[r] = fmaf(a.x, w.x, acc[r]); acc[r] = fmaf(a.y, w.y, acc[r]);
                    acc[r] = fmaf(a.z, w.z, acc[r]); acc[r] = fmaf(a.w, w.w, acc[r]);
                }
            }
        }
    }
    float mine = 0.f;
#pragma unroll
    for (int r = 0; r < SMALL_M; ++r) {
        float s = acc[r];
#pragma unroll
        for (int o = 16; o >= 1; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (lane == r) mine = s;
    }
    if (lane >= M) return;
    float v = mine;
    if (EPI == EPI_ACT) {
        const float a = v + p0[j];
        const float alpha = p1[j], beta = p2[j];
        const float s = 1.f / (1.f + expf(-alpha * a));      // exp(+big)=inf -> s=0, as numpy
        v = (beta + (1.f - beta) * s) * a;
    } else {
        v = fmaf(v, p0[j], p1[j]);
        if (EPI == EPI_OUT && accumulate) v += C[lane * ldc + j];      // fused sum of log-spectra
        if (EPI == EPI_OUT && ten_to) v = exp10f(v);
    }
    C[lane * ldc + j] = v;
}

}  // namespace cpb
