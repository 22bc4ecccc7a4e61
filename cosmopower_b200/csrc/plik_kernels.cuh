// Planck 2018 plik-lite likelihood tail behind the TT / TE / EE emulators:
// bin the three spectra into 613 band powers and evaluate chi^2 = d^T F d per parameter set.
//
// Math follows reference cosmopower/likelihoods/tf_planck2018_lite/tf_planck2018_lite.py:281-304
// (units factor, gather * window, segment sum, 1 / A_planck^2, difference to the data vector, Fisher
// quadratic form, -chi^2 / 2); the gather ranges are the reference's (:126-202), passed in as CSR tables.
//
// fp32 CUDA-core kernels (the reference computes this part in float32 as well):
//   plik_bin_kernel        HBM-bound: reads 3 x n_ell floats per row once (coalesced, staged in shared memory),
//                          writes the 640-float padded difference vector
//   plik_quadform_kernel   (n_rows x 640) x (640 x 640) with the dot product against d fused into the epilogue;
//                          F is symmetric, so only the upper-triangular 64-wide blocks are visited (F' = strict
//                          upper triangle + half the diagonal, chi^2 = 2 d^T F' d): 55 % of the dense flops
//   plik_finish_kernel     sums the 10 column-block partials per row (fixed order: deterministic)
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace cpb {

constexpr int PLIK_BIN_THREADS = 256;
constexpr int PLIK_QF_BM = 128, PLIK_QF_BN = 64, PLIK_QF_BK = 16;

// One block per row (grid-stride).  smem: the row's three spectra, concatenated as the reference does.
__global__ void __launch_bounds__(PLIK_BIN_THREADS)
plik_bin_kernel(const float* __restrict__ cl_tt, const float* __restrict__ cl_te, const float* __restrict__ cl_ee,
                long long pitch, const float* __restrict__ cal, long long n_rows, int n_ell,
                const int* __restrict__ bin_start, const int* __restrict__ gather_first,
                const float* __restrict__ window, const float* __restrict__ x_data, int n_bins, float units_factor,
                float* __restrict__ diff, int diff_pitch, float* __restrict__ binned, long long binned_pitch)
{
    extern __shared__ float row[];                 // 3 * n_ell
    const bool vec = (pitch & 3) == 0 &&
                     ((reinterpret_cast<uintptr_t>(cl_tt) | reinterpret_cast<uintptr_t>(cl_te) | reinterpret_cast<uintptr_t>(cl_ee)) & 15) == 0;
    for (long long r = blockIdx.x; r < n_rows; r += gridDim.x) {
        const float* src[3] = {cl_tt + r * pitch, cl_te + r * pitch, cl_ee + r * pitch};
        __syncthreads();                           // previous row's readers are done
#pragma unroll
        for (int s = 0; s < 3; ++s) {
            float* dst = row + s * n_ell;
            if (vec) {
                // rows start 16-byte aligned; the shared copy of spectrum s starts at s * n_ell floats, so go through
                // registers and store scalars (shared memory stores are cheap, global loads are the 16-byte ones)
                const int n4 = n_ell >> 2;
                for (int i = threadIdx.x; i < n4; i += PLIK_BIN_THREADS) {
                    const float4 v = __ldcs(reinterpret_cast<const float4*>(src[s]) + i);
                    dst[4 * i] = v.x; dst[4 * i + 1] = v.y; dst[4 * i + 2] = v.z; dst[4 * i + 3] = v.w;
                }
                for (int i = (n4 << 2) + threadIdx.x; i < n_ell; i += PLIK_BIN_THREADS) dst[i] = __ldcs(src[s] + i);
            } else {
                for (int i = threadIdx.x; i < n_ell; i += PLIK_BIN_THREADS) dst[i] = __ldcs(src[s] + i);
            }
        }
        __syncthreads();
        const float c = cal[r];
        const float scale = units_factor / (c * c);
        for (int b = threadIdx.x; b < diff_pitch; b += PLIK_BIN_THREADS) {
            float d = 0.f;
            if (b < n_bins) {
                const int j0 = bin_start[b], j1 = bin_start[b + 1];
                const float* p = row + gather_first[b];          // a bin gathers one contiguous run of multipoles
                float acc = 0.f;
                for (int j = j0; j < j1; ++j) acc = fmaf(p[j - j0], window[j], acc);
                const float model = acc * scale;
                if (binned) binned[r * binned_pitch + b] = model;
                d = x_data[b] - model;
            }
            diff[r * diff_pitch + b] = d;
        }
    }
}

// partial[r][jt] = sum over the 64 columns n of block jt of d[r][n] * (sum_{k <= last row of block jt} d[r][k] F'[k][n])
__global__ void __launch_bounds__(256)
plik_quadform_kernel(const float* __restrict__ D, int ldd, const float* __restrict__ Fp, int ldf, long long n_rows,
                     float* __restrict__ partial, int n_jt)
{
    __shared__ float As[PLIK_QF_BK][PLIK_QF_BM + 4];
    __shared__ float Bs[PLIK_QF_BK][PLIK_QF_BN + 4];
    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;            // 16 x 16 threads: 4 columns x 8 rows each
    const int jt = blockIdx.x;
    const long long row0 = static_cast<long long>(blockIdx.y) * PLIK_QF_BM;
    const int col0 = jt * PLIK_QF_BN;
    const int k_end = col0 + PLIK_QF_BN;               // F' is upper triangular: rows k beyond this block's columns are zero

    float acc[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    for (int k0 = 0; k0 < k_end; k0 += PLIK_QF_BK) {
        {   // A tile: 128 rows x 16 k = 512 float4, two per thread
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                const int f = tid + u * 256, r = f >> 2, kq = (f & 3) * 4;
                const long long gr = row0 + r;
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (gr < n_rows) v = *reinterpret_cast<const float4*>(D + gr * ldd + k0 + kq);
                As[kq][r] = v.x; As[kq + 1][r] = v.y; As[kq + 2][r] = v.z; As[kq + 3][r] = v.w;
            }
        }
        {   // B tile: 16 k x 64 n = 256 float4, one per thread
            const int kk = tid >> 4, nq = (tid & 15) * 4;
            *reinterpret_cast<float4*>(&Bs[kk][nq]) = *reinterpret_cast<const float4*>(Fp + static_cast<long long>(k0 + kk) * ldf + col0 + nq);
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < PLIK_QF_BK; ++kk) {
            const float4 a0 = *reinterpret_cast<const float4*>(&As[kk][ty * 8]);
            const float4 a1 = *reinterpret_cast<const float4*>(&As[kk][ty * 8 + 4]);
            const float4 b = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
            const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
            const float w[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], w[j], acc[i][j]);
        }
        __syncthreads();
    }
    // epilogue: dot with this block's columns of d, reduce over the 16 threads that share a row
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const long long gr = row0 + ty * 8 + i;
        float s = 0.f;
        if (gr < n_rows) {
            const float4 d = *reinterpret_cast<const float4*>(D + gr * ldd + col0 + tx * 4);
            s = acc[i][0] * d.x + acc[i][1] * d.y + acc[i][2] * d.z + acc[i][3] * d.w;
        }
#pragma unroll
        for (int o = 8; o >= 1; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o, 16);
        if (tx == 0 && gr < n_rows) partial[gr * n_jt + jt] = s;
    }
}

// loglike = -0.5 * chi^2 = -(sum of the partials), since chi^2 = 2 d^T F' d
__global__ void plik_finish_kernel(const float* __restrict__ partial, int n_jt, long long n_rows, float* __restrict__ loglike)
{
    const long long r = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (r >= n_rows) return;
    float s = 0.f;
    for (int j = 0; j < n_jt; ++j) s += partial[r * n_jt + j];
    loglike[r] = -s;
}

}  // namespace cpb
