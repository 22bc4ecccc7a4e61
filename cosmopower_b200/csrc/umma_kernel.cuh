// Persistent fused-MLP kernel for sm_100a (CPB200_PREC_F16X3).
//
// One CTA per SM walks pairs of 128-row tiles through the WHOLE network (standardisation, every
// dense layer, activation, PCA rescale + basis contraction, de-standardisation, 10**x).
// Every dense contraction runs on the 5th-gen tensor cores (tcgen05.mma kind::f16, M=128,
// N<=256, fp32 accumulators in TMEM) as a 3-product fp16 split:
//     A.W  ~=  A_hi.W_hi + A_lo.W_hi + A_hi.W_lo      (x_hi = fp16(x), x_lo = fp16(x - x_hi))
// which keeps ~22 significand bits per operand (single-pass bf16/tf32 miss the 1e-5 tolerance,
// SURVEY 7.2).
//
// Data movement: operands are kept in the tensor core's canonical no-swizzle K-major
// "core matrix" layout (8 rows x 16 bytes, contiguous 128 B) both in global memory and in shared
// memory, so one pipeline stage is two plain bulk copies (cp.async.bulk, the TMA engine's 1-D
// mode): 16 KB of activations (hi|lo, 128 rows x 32 k) and 2*NC*64 B of weights (hi|lo).
// Activations between layers go through a CTA-private scratch that lives in L2 (a 128x512 tile in
// split precision is 256 KB: more than shared memory, and TMEM holds the accumulators);
// the epilogue warps write it with full-line 16-byte stores and the loader streams it back.
// Two row tiles are interleaved per CTA so the epilogue of one overlaps the MMAs of the other.
//
// Warp roles (384 threads): warp 0 loader (bulk copies), warp 1 MMA issuer (one thread),
// warp 2 TMEM allocator + parameter standardiser, warp 3 idle, warps 4..11 epilogue.
//
// Reference math: cosmopower/cosmopower_NN.py:371-384,425 and
// cosmopower/cosmopower_PCAplusNN.py:368-381,421.
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "simt_kernels.cuh"   // Epilogue enum

namespace cpb {

constexpr int UM_TILE_M = 128;
constexpr int UM_KS = 32;                                    // k elements per pipeline stage
constexpr int UM_MAX_NC = 256;                               // widest N chunk (one accumulator)
constexpr int UM_STAGES = 4;
constexpr int UM_A_HALF_BYTES = UM_TILE_M * UM_KS * 2;       // 8192: one of (hi, lo)
constexpr int UM_A_STAGE_BYTES = 2 * UM_A_HALF_BYTES;        // 16384
constexpr int UM_B_STAGE_BYTES_MAX = 2 * UM_MAX_NC * UM_KS * 2;   // 32768
constexpr int UM_STAGE_BYTES = UM_A_STAGE_BYTES + UM_B_STAGE_BYTES_MAX;   // 49152
constexpr int UM_MAX_GEMMS = 8;
constexpr int UM_THREADS = 384;
constexpr int UM_EPI_WARP0 = 4;
constexpr int UM_EPI_WARPS = 8;
constexpr int UM_EPI_THREADS = UM_EPI_WARPS * 32;
constexpr int UM_PIECE = 16;                                 // accumulator columns per tcgen05.ld
constexpr float UM_INPUT_SCALE = 1024.f;                     // power-of-two pre-scale of the standardised parameters
constexpr float UM_ACT_SCALE = 64.f;                         // power-of-two pre-scale of every produced activation

// shared memory carve-up (bytes from the 1024-aligned base)
constexpr int UM_SMEM_STAGES = 0;
constexpr int UM_SMEM_OUTSTG = UM_STAGES * UM_STAGE_BYTES;                   // 8 warps x 32x16 floats
constexpr int UM_SMEM_EPIPAR = UM_SMEM_OUTSTG + UM_EPI_WARPS * 32 * 16 * 4;  // 2 x 256 float4
constexpr int UM_SMEM_BARS = UM_SMEM_EPIPAR + 2 * UM_MAX_NC * 16;
constexpr int UM_NUM_BARS = 2 * UM_STAGES + 4 + 2 + 2 + 2;
constexpr int UM_SMEM_TMEMPTR = UM_SMEM_BARS + UM_NUM_BARS * 8;
constexpr int UM_SMEM_BYTES = UM_SMEM_TMEMPTR + 16 + 1024;                   // + alignment slack

struct UmmaGemm {
    const __half* wpack;   // [n_chunks][k_stages][hi|lo][nc x 32] core-matrix tiles
    const float4* epi;     // per padded output column: ACT {b, -alpha*log2e, beta*s, (1-beta)*s}; else {scale, shift,0,0}
    int k_stages;          // K_pad / 32
    int nc;                // chunk width: multiple of 32, <= 256
    int n_chunks;
    int epi_kind;          // cpb::Epilogue
    float acc_scale;       // undoes the power-of-two operand pre-scales on the accumulator (EPI_ACT)
};

struct UmmaParams {
    UmmaGemm g[UM_MAX_GEMMS];
    int n_gemms;
    int n_parameters;
    int n_modes;
    int ten_to;
    const float* params;       // (n_rows, n_parameters)
    const float* pmean;
    const float* pstd;
    float* out;
    long long out_pitch;
    long long n_rows;
    long long n_pairs;         // ceil(n_tiles / 2)
    uint8_t* scratch;          // gridDim.x * scratch_cta_bytes
    unsigned long long scratch_cta_bytes;
    unsigned a0_bytes;         // per tile: k_stages(gemm 0) * 16384
    unsigned abuf_bytes;       // per tile per buffer: max k_stages(gemm >= 1) * 16384
    int* error_flag;           // set to a code by the watchdog before trapping
    long long* counters;       // optional debug counters, 16 per CTA (nullptr = off); see UM_CNT_*
};

// debug counter slots (cycles), per CTA
enum { UM_CNT_LOAD_WAIT_EMPTY = 0, UM_CNT_LOAD_WAIT_AREADY = 1, UM_CNT_LOAD_TOTAL = 2,
       UM_CNT_MMA_WAIT_FULL = 3, UM_CNT_MMA_WAIT_ACCEMPTY = 4, UM_CNT_MMA_TOTAL = 5,
       UM_CNT_EPI_WAIT_ACCFULL = 6, UM_CNT_EPI_ACT = 7, UM_CNT_EPI_OUT = 8, UM_CNT_EPI_TOTAL = 9,
       UM_CNT_IN_WAIT = 10, UM_CNT_IN_TOTAL = 11, UM_CNT_SLOTS = 16 };

// ---------------------------------------------------------------------------------------------
// PTX wrappers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ uint32_t mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok;
}
#ifndef CPB_WATCHDOG_CYCLES
#define CPB_WATCHDOG_CYCLES 6000000000LL   // ~3 s at 2 GHz: a stuck pipeline traps instead of hanging the GPU
#endif
__device__ __noinline__ long long mbar_wait_slow(uint32_t bar, uint32_t parity, int* error_flag, int code) {
    const long long t0 = clock64();
    long long t1 = t0;
    while (!mbar_try_wait(bar, parity)) {
        t1 = clock64();
        if (t1 - t0 > CPB_WATCHDOG_CYCLES) {
            if (error_flag) atomicExch(error_flag, code);
            __threadfence_system();
            __trap();
        }
    }
    return clock64() - t0;
}
// returns the cycles spent blocked (0 on the fast path); callers may add it to a debug counter
__device__ __forceinline__ long long mbar_wait(uint32_t bar, uint32_t parity, int* error_flag, int code) {
    if (mbar_try_wait(bar, parity)) return 0;
    return mbar_wait_slow(bar, parity, error_flag, code);
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst_smem, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst_smem), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void umma_f16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                         uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// K-major, no-swizzle shared-memory matrix descriptor: core matrices (8 rows x 16 B) are 128 B
// apart along K (LBO) and 512 B apart along M/N (SBO = 4 core matrices of a 32-wide stage).
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr) {
    uint64_t d = 0;
    d |= static_cast<uint64_t>((saddr & 0x3FFFFu) >> 4);
    d |= static_cast<uint64_t>(128u >> 4) << 16;
    d |= static_cast<uint64_t>(512u >> 4) << 32;
    d |= static_cast<uint64_t>(1) << 46;              // descriptor version for sm_100
    return d;
}
// kind::f16 instruction descriptor: D=f32, A=B=f16, both K-major, M=128, N=nc
__device__ __forceinline__ uint32_t umma_idesc(int nc) {
    return (1u << 4) | (static_cast<uint32_t>(nc >> 3) << 17) | (static_cast<uint32_t>(UM_TILE_M >> 4) << 24);
}

__device__ __forceinline__ uint32_t pack_half2(float a, float b) {
    __half2 h = __floats2half2_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
}
// split two floats into fp16 hi and lo parts (hi = rn(x), lo = rn(x - hi)), packed as half2
__device__ __forceinline__ void split2(float a, float b, uint32_t& hi, uint32_t& lo) {
    __half2 h = __floats2half2_rn(a, b);
    float2 hf = __half22float2(h);
    __half2 l = __floats2half2_rn(a - hf.x, b - hf.y);
    hi = *reinterpret_cast<uint32_t*>(&h);
    lo = *reinterpret_cast<uint32_t*>(&l);
}
__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float rcp_approx(float x) {
    float y;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// ---------------------------------------------------------------------------------------------
// the kernel
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(UM_THREADS, 1)
fused_mlp_umma_kernel(const __grid_constant__ UmmaParams P)
{
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;
    uint8_t* sbase = smem_raw + (base - raw);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    const uint32_t bars = base + UM_SMEM_BARS;
    auto bar_full = [&](int s) { return bars + 8u * s; };
    auto bar_empty = [&](int s) { return bars + 8u * (UM_STAGES + s); };
    auto bar_accfull = [&](int b) { return bars + 8u * (2 * UM_STAGES + b); };
    auto bar_accempty = [&](int b) { return bars + 8u * (2 * UM_STAGES + 2 + b); };
    auto bar_aready = [&](int t) { return bars + 8u * (2 * UM_STAGES + 4 + t); };
    auto bar_a0ready = [&](int t) { return bars + 8u * (2 * UM_STAGES + 6 + t); };
    auto bar_a0free = [&](int t) { return bars + 8u * (2 * UM_STAGES + 8 + t); };
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(sbase + UM_SMEM_TMEMPTR);

    if (threadIdx.x == 0) {
        for (int s = 0; s < UM_STAGES; ++s) { mbar_init(bar_full(s), 1); mbar_init(bar_empty(s), 1); }
        for (int b = 0; b < 2; ++b) { mbar_init(bar_accfull(b), 1); mbar_init(bar_accempty(b), UM_EPI_WARPS); }
        for (int t = 0; t < 2; ++t) {
            mbar_init(bar_aready(t), UM_EPI_WARPS);
            mbar_init(bar_a0ready(t), 1);
            mbar_init(bar_a0free(t), 1);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                     ::"r"(smem_u32(tmem_ptr_smem)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;

    uint8_t* cta_scratch = P.scratch + static_cast<unsigned long long>(blockIdx.x) * P.scratch_cta_bytes;
    auto a0_buf = [&](int t) { return cta_scratch + static_cast<size_t>(t) * P.a0_bytes; };
    auto a_buf = [&](int t, int b) {
        return cta_scratch + 2ull * P.a0_bytes + (static_cast<size_t>(t) * 2 + b) * P.abuf_bytes;
    };
    const int G = P.n_gemms;

    if (warp == 0) {
        // ===================== loader: bulk copies into the stage ring =====================
        if (lane == 0) {
            long long w_empty = 0, w_aready = 0; const long long t_begin = clock64();
            int stage = 0; uint32_t ph_empty = 1;                 // fresh "empty" barriers pass at parity 1
            uint32_t ph_aready[2] = {0, 0}, ph_a0ready[2] = {0, 0};
            for (long long pair = blockIdx.x; pair < P.n_pairs; pair += gridDim.x) {
                for (int g = 0; g < G; ++g) {
                    const UmmaGemm& L = P.g[g];
                    const uint32_t b_bytes = 2u * L.nc * UM_KS * 2u;
                    for (int t = 0; t < 2; ++t) {
                        const uint8_t* a_src;
                        if (g == 0) {
                            w_aready += mbar_wait(bar_a0ready(t), ph_a0ready[t], P.error_flag, 101); ph_a0ready[t] ^= 1;
                            a_src = a0_buf(t);
                        } else {
                            w_aready += mbar_wait(bar_aready(t), ph_aready[t], P.error_flag, 102); ph_aready[t] ^= 1;
                            a_src = a_buf(t, g & 1);
                        }
                        for (int c = 0; c < L.n_chunks; ++c) {
                            const uint8_t* b_src = reinterpret_cast<const uint8_t*>(L.wpack)
                                                   + static_cast<size_t>(c) * L.k_stages * b_bytes;
                            for (int s = 0; s < L.k_stages; ++s) {
                                w_empty += mbar_wait(bar_empty(stage), ph_empty, P.error_flag, 103);
                                const uint32_t dst = base + UM_SMEM_STAGES + stage * UM_STAGE_BYTES;
                                mbar_arrive_expect_tx(bar_full(stage), UM_A_STAGE_BYTES + b_bytes);
                                bulk_g2s(dst, a_src + static_cast<size_t>(s) * UM_A_STAGE_BYTES, UM_A_STAGE_BYTES, bar_full(stage));
                                bulk_g2s(dst + UM_A_STAGE_BYTES, b_src + static_cast<size_t>(s) * b_bytes, b_bytes, bar_full(stage));
                                if (++stage == UM_STAGES) { stage = 0; ph_empty ^= 1; }
                            }
                        }
                    }
                }
            }
            if (P.counters) {
                long long* cnt = P.counters + blockIdx.x * UM_CNT_SLOTS;
                cnt[UM_CNT_LOAD_WAIT_EMPTY] = w_empty; cnt[UM_CNT_LOAD_WAIT_AREADY] = w_aready;
                cnt[UM_CNT_LOAD_TOTAL] = clock64() - t_begin;
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        if (lane == 0) {
            long long w_full = 0, w_accempty = 0; const long long t_begin = clock64();
            int stage = 0; uint32_t ph_full = 0;
            uint32_t ph_accempty[2] = {1, 1};
            int job = 0;
            for (long long pair = blockIdx.x; pair < P.n_pairs; pair += gridDim.x) {
                for (int g = 0; g < G; ++g) {
                    const UmmaGemm& L = P.g[g];
                    const uint32_t idesc = umma_idesc(L.nc);
                    const uint32_t b_half = static_cast<uint32_t>(L.nc) * UM_KS * 2u;
                    for (int t = 0; t < 2; ++t) {
                        for (int c = 0; c < L.n_chunks; ++c, ++job) {
                            const int buf = job & 1;
                            w_accempty += mbar_wait(bar_accempty(buf), ph_accempty[buf], P.error_flag, 201); ph_accempty[buf] ^= 1;
                            tc_fence_after();
                            const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(buf) * UM_MAX_NC;
                            for (int s = 0; s < L.k_stages; ++s) {
                                w_full += mbar_wait(bar_full(stage), ph_full, P.error_flag, 202);
                                tc_fence_after();
                                const uint32_t sa = base + UM_SMEM_STAGES + stage * UM_STAGE_BYTES;
                                const uint64_t a_hi = umma_desc(sa);
                                const uint64_t a_lo = umma_desc(sa + UM_A_HALF_BYTES);
                                const uint64_t b_hi = umma_desc(sa + UM_A_STAGE_BYTES);
                                const uint64_t b_lo = umma_desc(sa + UM_A_STAGE_BYTES + b_half);
#pragma unroll
                                for (int kk = 0; kk < UM_KS / 16; ++kk) {
                                    const uint64_t adv = static_cast<uint64_t>((kk * 256) >> 4);   // 2 core matrices along K
                                    umma_f16(d_tmem, a_hi + adv, b_hi + adv, idesc, (s | kk) != 0);
                                    umma_f16(d_tmem, a_lo + adv, b_hi + adv, idesc, 1u);
                                    umma_f16(d_tmem, a_hi + adv, b_lo + adv, idesc, 1u);
                                }
                                umma_commit(bar_empty(stage));
                                if (++stage == UM_STAGES) { stage = 0; ph_full ^= 1; }
                            }
                            umma_commit(bar_accfull(buf));
                        }
                        if (g == 0) umma_commit(bar_a0free(t));
                    }
                }
            }
            if (P.counters) {
                long long* cnt = P.counters + blockIdx.x * UM_CNT_SLOTS;
                cnt[UM_CNT_MMA_WAIT_FULL] = w_full; cnt[UM_CNT_MMA_WAIT_ACCEMPTY] = w_accempty;
                cnt[UM_CNT_MMA_TOTAL] = clock64() - t_begin;
            }
        }
    } else if (warp == 2) {
        // ===================== parameter standardiser: builds the layer-0 A operand =====================
        uint32_t ph_a0free[2] = {1, 1};
        long long w_in = 0; const long long t_begin = clock64();
        const int np = P.n_parameters;
        for (long long pair = blockIdx.x; pair < P.n_pairs; pair += gridDim.x) {
            for (int t = 0; t < 2; ++t) {
                w_in += mbar_wait(bar_a0free(t), ph_a0free[t], P.error_flag, 301); ph_a0free[t] ^= 1;
                const long long row0 = (pair * 2 + t) * UM_TILE_M;
                uint8_t* dst = a0_buf(t);
                for (int rr = 0; rr < UM_TILE_M / 32; ++rr) {
                    const int r = rr * 32 + lane;
                    const long long grow = row0 + r;
                    for (int k0 = 0; k0 < np; k0 += 8) {
                        float x[8];
#pragma unroll
                        for (int u = 0; u < 8; ++u) {
                            const int k = k0 + u;
                            float v = 0.f;
                            if (k < np) {
                                const float raw_v = (grow < P.n_rows) ? P.params[grow * np + k] : P.pmean[k];
                                v = (raw_v - P.pmean[k]) / P.pstd[k] * UM_INPUT_SCALE;
                            }
                            x[u] = v;
                        }
                        uint4 hi, lo;
                        split2(x[0], x[1], hi.x, lo.x); split2(x[2], x[3], hi.y, lo.y);
                        split2(x[4], x[5], hi.z, lo.z); split2(x[6], x[7], hi.w, lo.w);
                        const size_t off = static_cast<size_t>(k0 / UM_KS) * UM_A_STAGE_BYTES
                                           + (r >> 3) * 512 + ((k0 % UM_KS) >> 3) * 128 + (r & 7) * 16;
                        *reinterpret_cast<uint4*>(dst + off) = hi;
                        *reinterpret_cast<uint4*>(dst + off + UM_A_HALF_BYTES) = lo;
                    }
                }
                __threadfence();
                fence_proxy_async();
                __syncwarp();
                if (lane == 0) mbar_arrive(bar_a0ready(t));
            }
        }
        if (P.counters && lane == 0) {
            long long* cnt = P.counters + blockIdx.x * UM_CNT_SLOTS;
            cnt[UM_CNT_IN_WAIT] = w_in; cnt[UM_CNT_IN_TOTAL] = clock64() - t_begin;
        }
    } else if (warp >= UM_EPI_WARP0) {
        // ===================== epilogue warps =====================
        const int ew = warp - UM_EPI_WARP0;          // 0..7
        const int q = warp & 3;                      // TMEM lane quarter this warp may read
        const int hf = ew >> 2;                      // which half of the chunk's columns
        const int r = q * 32 + lane;                 // tile row owned by this thread
        const int et = threadIdx.x - UM_EPI_WARP0 * 32;
        float4* epipar = reinterpret_cast<float4*>(sbase + UM_SMEM_EPIPAR);
        float* outstg = reinterpret_cast<float*>(sbase + UM_SMEM_OUTSTG) + ew * (32 * 16);
        uint32_t ph_accfull[2] = {0, 0};
        long long w_accfull = 0, c_act = 0, c_out = 0; const long long t_begin = clock64();
        int job = 0;
        for (long long pair = blockIdx.x; pair < P.n_pairs; pair += gridDim.x) {
            for (int g = 0; g < G; ++g) {
                const UmmaGemm& L = P.g[g];
                const int half_cols = L.nc >> 1;
                for (int t = 0; t < 2; ++t) {
                    const long long row0 = (pair * 2 + t) * UM_TILE_M;
                    uint8_t* a_dst = a_buf(t, (g + 1) & 1);
                    for (int c = 0; c < L.n_chunks; ++c, ++job) {
                        const int buf = job & 1;
                        // stage this chunk's per-column epilogue constants (double-buffered by job parity)
                        float4* par = epipar + buf * UM_MAX_NC;
                        if (et < L.nc) par[et] = L.epi[c * L.nc + et];
                        asm volatile("bar.sync 1, %0;" ::"n"(UM_EPI_THREADS) : "memory");
                        w_accfull += mbar_wait(bar_accfull(buf), ph_accfull[buf], P.error_flag, 401); ph_accfull[buf] ^= 1;
                        tc_fence_after();
                        const long long t_job = clock64();
                        const uint32_t t_addr = tmem_base + (static_cast<uint32_t>(q * 32) << 16)
                                                + static_cast<uint32_t>(buf * UM_MAX_NC + hf * half_cols);
                        for (int pc = 0; pc < half_cols; pc += UM_PIECE) {
                            uint32_t v[16];
                            tmem_ld16(t_addr + pc, v);
                            tmem_ld_wait();
                            const int ccol = hf * half_cols + pc;          // column within the chunk
                            if (L.epi_kind == EPI_OUT) {
                                // transpose 32 rows x 16 cols through shared memory, then row-contiguous stores
                                float4* srow = reinterpret_cast<float4*>(outstg + lane * 16);
                                const int sw = (lane >> 1) & 3;
#pragma unroll
                                for (int j = 0; j < 4; ++j)
                                    srow[j ^ sw] = make_float4(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]),
                                                               __uint_as_float(v[4 * j + 2]), __uint_as_float(v[4 * j + 3]));
                                __syncwarp();
                                const int cl = lane & 15;
                                const float4 pp = par[ccol + cl];
                                const long long gcol = static_cast<long long>(c) * L.nc + ccol + cl;
#pragma unroll 4
                                for (int it = 0; it < 16; ++it) {
                                    const int rl = 2 * it + (lane >> 4);
                                    const float a = outstg[rl * 16 + (((cl >> 2) ^ ((rl >> 1) & 3)) << 2) + (cl & 3)];
                                    float y = fmaf(a, pp.x, pp.y);
                                    if (P.ten_to) y = ex2_approx(y * 3.3219280948873623f);
                                    const long long grow = row0 + q * 32 + rl;
                                    if (grow < P.n_rows && gcol < P.n_modes)
                                        __stcs(P.out + grow * P.out_pitch + gcol, y);
                                }
                                __syncwarp();
                            } else {
                                float h[16];
                                if (L.epi_kind == EPI_ACT) {
#pragma unroll
                                    for (int j = 0; j < 16; ++j) {
                                        const float4 pp = par[ccol + j];
                                        const float a = fmaf(__uint_as_float(v[j]), L.acc_scale, pp.x);
                                        const float e = ex2_approx(a * pp.y);          // exp(-alpha*a)
                                        const float s = rcp_approx(1.f + e);           // inf -> 0, as numpy
                                        h[j] = fmaf(pp.w, s, pp.z) * a;
                                    }
                                } else {
#pragma unroll
                                    for (int j = 0; j < 16; ++j) {
                                        const float4 pp = par[ccol + j];
                                        h[j] = fmaf(__uint_as_float(v[j]), pp.x, pp.y);
                                    }
                                }
                                const int k = c * L.nc + ccol;                         // next layer's k index
#pragma unroll
                                for (int u = 0; u < 2; ++u) {
                                    uint4 hi, lo;
                                    split2(h[8 * u + 0], h[8 * u + 1], hi.x, lo.x);
                                    split2(h[8 * u + 2], h[8 * u + 3], hi.y, lo.y);
                                    split2(h[8 * u + 4], h[8 * u + 5], hi.z, lo.z);
                                    split2(h[8 * u + 6], h[8 * u + 7], hi.w, lo.w);
                                    const int kk = k + 8 * u;
                                    const size_t off = static_cast<size_t>(kk / UM_KS) * UM_A_STAGE_BYTES
                                                       + (r >> 3) * 512 + ((kk % UM_KS) >> 3) * 128 + (r & 7) * 16;
                                    *reinterpret_cast<uint4*>(a_dst + off) = hi;
                                    *reinterpret_cast<uint4*>(a_dst + off + UM_A_HALF_BYTES) = lo;
                                }
                            }
                        }
                        // accumulator drained: hand the TMEM buffer back to the MMA issuer
                        tc_fence_before();
                        if (L.epi_kind != EPI_OUT) { __threadfence(); fence_proxy_async(); }
                        __syncwarp();
                        if (lane == 0) {
                            mbar_arrive(bar_accempty(buf));
                            if (L.epi_kind != EPI_OUT && c == L.n_chunks - 1) mbar_arrive(bar_aready(t));
                        }
                        if (L.epi_kind == EPI_OUT) c_out += clock64() - t_job; else c_act += clock64() - t_job;
                    }
                }
            }
        }
        if (P.counters && ew == 0 && lane == 0) {
            long long* cnt = P.counters + blockIdx.x * UM_CNT_SLOTS;
            cnt[UM_CNT_EPI_WAIT_ACCFULL] = w_accfull; cnt[UM_CNT_EPI_ACT] = c_act; cnt[UM_CNT_EPI_OUT] = c_out;
            cnt[UM_CNT_EPI_TOTAL] = clock64() - t_begin;
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    }
}

}  // namespace cpb
