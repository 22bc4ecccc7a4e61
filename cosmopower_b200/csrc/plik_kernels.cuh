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
//                          F is symmetric, so only the upper-triangular 128-wide blocks are visited (F' = strict
//                          upper triangle + half the diagonal, chi^2 = 2 d^T F' d): 60 % of the dense flops
//   plik_finish_kernel     sums the 5 column-block partials per row (fixed order: deterministic)
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <cuda_fp16.h>

namespace cpb {

constexpr int PLIK_BIN_THREADS = 256;
constexpr int PLIK_QF_BM = 128, PLIK_QF_BN = 128, PLIK_QF_BK = 16;

// Persistent blocks (three per SM) walking rows through a shared-memory ring of PLIK_GROUPS groups of PLIK_R row buffers: while a group is being binned
// the next one is already on its way (one bulk copy per input and row, completion on the group's mbarrier;
// each input starts 16-byte aligned in shared memory).  A slot is a run of at most PLIK_SEG consecutive multipoles of one bin; the host splits the 5 ... 33
// wide bins into slots so that the per-thread work is short, balanced and fully unrolled, and lays a bin's slots out on
// neighbouring lanes of one warp.  Every thread owns the same <= PLIK_MAX_PASSES slots for every row, so their window
// weights live in registers; the lane holding a bin's first slot adds the other partial sums (at most
// PLIK_MAX_SLOTS_PER_BIN - 1 shuffles, fixed order: deterministic) and writes d = X_data - model.
constexpr int PLIK_SEG = 9;
constexpr int PLIK_MAX_PASSES = 3;
constexpr int PLIK_MAX_SLOTS_PER_BIN = 4;
constexpr int PLIK_R = 2;                // rows binned per loop iteration (per block barrier)
constexpr int PLIK_GROUPS = 2;           // ring of row-buffer groups per block: one group being binned, one in flight
constexpr int PLIK_BLOCKS_PER_SM = 3;      // 80 registers per thread, 3 x 63 KB of shared memory

__device__ __forceinline__ void cp_async_16(void* smem_dst, const void* gsrc) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(static_cast<uint32_t>(__cvta_generic_to_shared(smem_dst))), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_4(void* smem_dst, const void* gsrc) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(static_cast<uint32_t>(__cvta_generic_to_shared(smem_dst))), "l"(gsrc) : "memory");
}

// fp16 hi / -lo split of two floats (same values as the fused kernel's split2n: its A operand tiles hold the lo part
// negated, the tensor core's negate-A bit takes the sign back)
__device__ __forceinline__ void plik_split2(float a, float b, uint32_t& hi, uint32_t& lo) {
    const __half2 h = __floats2half2_rn(a, b);
    const float2 hf = __half22float2(h);
    const __half2 l = __floats2half2_rn(hf.x - a, hf.y - b);
    hi = *reinterpret_cast<const uint32_t*>(&h);
    lo = *reinterpret_cast<const uint32_t*>(&l);
}

struct PlikSlot {          // one per slot, built by cpb200_plik_create
    int off;               // shared-memory offset of the slot's first element (input s starts 16-byte aligned after input s-1)
    int last;              // index of the slot's last multipole (the kernel always reads PLIK_SEG elements; weights beyond are 0)
    int bin;               // the bin this slot opens, or -1
    int n_slots;           // slots of that bin (1 .. PLIK_MAX_SLOTS_PER_BIN) when bin >= 0
    float x_data;          // data band power of that bin
    float sqrt_f;          // sqrt(F[bin][bin]): scales the difference for the tensor-core quadratic form
    float w[PLIK_SEG];     // window weights, zero beyond the slot's end
};

// The three inputs of a row: TT, TE, EE.  Each is either a spectrum of n_ell multipoles or - TE, when its emulator was
// built with the binning folded into its (linear) PCA basis - that spectrum's band powers already.
struct PlikInputs {
    const float* ptr[3];
    long long pitch[3];
    int len[3];            // floats per row
};

// Shared memory of one block: the groups' mbarriers (32 bytes), the row ring - each buffer holds the three inputs back to
// back (every one starting 16-byte aligned) plus PLIK_ROW_PAD zero floats, because a slot always reads PLIK_SEG
// elements (those beyond its end carry weight 0) -, the row's scaled differences and the per-warp maxima.
constexpr int PLIK_ROW_PAD = 8;
constexpr int PLIK_BAR_FLOATS = 8;

// TILES: tensor-core quadratic form behind it (scaled differences + operand tiles); false: plain differences for the fp32 one
//
// What bounds this kernel is instruction issue and the block barrier, not HBM (ncu, round 2: 65 % of the issue slots and 490
// instructions per thread and row in its first form; 30 % of the warp time in the barrier once those were halved).  Hence:
// PLIK_R rows per loop iteration and barrier (the window weights in a thread's registers serve both), a row written out one
// iteration after it was binned (so that ONE barrier per iteration separates the ring hand-over and the row-maximum
// exchange), the fetch left to warp 0, the row / tile stores to the 80 threads that own a 16-byte chunk, shared-memory
// addresses kept as pinned 32-bit integers (the compiler otherwise re-derives every loop invariant from the parameters).
template <bool TILES>
__global__ void __launch_bounds__(PLIK_BIN_THREADS, PLIK_BLOCKS_PER_SM)
plik_bin_kernel(PlikInputs in, const float* __restrict__ cal, long long n_rows,
                const PlikSlot* __restrict__ slots, int n_passes, float units_factor,
                float* __restrict__ diff, int diff_pitch, float* __restrict__ binned, long long binned_pitch,
                uint8_t* __restrict__ tiles, float* __restrict__ row_scale)
{
    extern __shared__ __align__(16) float plik_smem[];
    constexpr int R = PLIK_R, G = PLIK_GROUPS;
    int off1 = (in.len[0] + 3) & ~3, off2 = off1 + ((in.len[1] + 3) & ~3);
    int buf_floats = off2 + ((in.len[2] + 3) & ~3) + PLIK_ROW_PAD;
    asm volatile("" : "+r"(off1), "+r"(off2), "+r"(buf_floats));      // pinned in registers
    // shared memory: G mbarriers | ring of G groups x R row buffers | 2 x R copies of {scaled differences, per-warp maxima}
    float* ring = plik_smem + PLIK_BAR_FLOATS;
    float* vrow0 = ring + G * R * buf_floats;
    const int vstride = diff_pitch + PLIK_BIN_THREADS / 32;
    const uint32_t bar0 = static_cast<uint32_t>(__cvta_generic_to_shared(plik_smem));
    // zero everything once: the pads between and after the inputs stay zero for the whole launch
    for (int i = threadIdx.x; i < G * R * buf_floats + 2 * R * vstride; i += PLIK_BIN_THREADS) ring[i] = 0.f;
    if (threadIdx.x == 0) {
        for (int i = 0; i < G; ++i) mbar_init(bar0 + 8u * i, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // the zero fill is ordered before the bulk copies' writes
    __syncthreads();

    // An input whose rows are 16-byte aligned arrives by one bulk copy per row (issued by thread 0, completion counted on
    // the group's mbarrier) plus at most three 4-byte cp.async for the elements beyond the last multiple of four; any
    // other input is gathered 4 bytes at a time by the whole block.
    bool vec[3];
    uint32_t bulk_bytes = 0;
#pragma unroll
    for (int s = 0; s < 3; ++s) {
        vec[s] = (in.pitch[s] & 3) == 0 && (reinterpret_cast<uintptr_t>(in.ptr[s]) & 15) == 0;
        if (vec[s]) bulk_bytes += static_cast<uint32_t>(in.len[s] >> 2) * 16u;
    }
    int all_vec = vec[0] && vec[1] && vec[2];
    asm volatile("" : "+r"(all_vec), "+r"(bulk_bytes));

    PlikSlot my[PLIK_MAX_PASSES];
#pragma unroll
    for (int p = 0; p < PLIK_MAX_PASSES; ++p) {
        if (p < n_passes) my[p] = slots[p * PLIK_BIN_THREADS + threadIdx.x];
        else { my[p].off = 0; my[p].last = 0; my[p].bin = -1; my[p].n_slots = 0; my[p].x_data = 0.f; my[p].sqrt_f = 0.f;
#pragma unroll
               for (int j = 0; j < PLIK_SEG; ++j) my[p].w[j] = 0.f; }
    }

    const long long stride = gridDim.x;
    // rows r, r + stride, ... r + (R - 1) stride into the R buffers of group g; always commits a cp.async group
    auto fetch = [&](long long r, int g) {
        const uint32_t bar = bar0 + 8u * g;
        int n_valid = 0;
#pragma unroll
        for (int j = 0; j < R; ++j) n_valid += (r + j * stride < n_rows) ? 1 : 0;
        if (all_vec) {
            // the LAST warp alone (the first ones carry the row write-out): its lane 0 issues the bulk copies, lanes
            // 1 + 10 j + (0 .. 8) the (at most three per input) tail elements of row j
            if (threadIdx.x < PLIK_BIN_THREADS - 32) return;
            const int lane = static_cast<int>(threadIdx.x) - (PLIK_BIN_THREADS - 32);
            if (lane == 0) {
                if (n_valid) mbar_arrive_expect_tx(bar, bulk_bytes * n_valid);
#pragma unroll
                for (int j = 0; j < R; ++j) {
                    const long long rj = r + j * stride;
                    if (rj < n_rows) {
                        float* dst0 = ring + (g * R + j) * buf_floats;
#pragma unroll
                        for (int s = 0; s < 3; ++s) {
                            const int n4 = in.len[s] >> 2;
                            if (n4) bulk_g2s(static_cast<uint32_t>(__cvta_generic_to_shared(dst0 + (s == 0 ? 0 : s == 1 ? off1 : off2))),
                                             in.ptr[s] + rj * in.pitch[s], static_cast<uint32_t>(n4) * 16u, bar);
                        }
                    }
                }
            } else {
                const int l = lane - 1, j = l / 10, k = l % 10;
                const long long rj = r + j * stride;
                if (j < R && k < 9 && rj < n_rows) {
                    const int s = k / 3, u = k % 3;
                    const int len = s == 0 ? in.len[0] : s == 1 ? in.len[1] : in.len[2];
                    const int i = (len & ~3) + u;
                    if (i < len) {
                        const float* src = (s == 0 ? in.ptr[0] : s == 1 ? in.ptr[1] : in.ptr[2]) + rj * (s == 0 ? in.pitch[0] : s == 1 ? in.pitch[1] : in.pitch[2]);
                        cp_async_4(ring + (g * R + j) * buf_floats + (s == 0 ? 0 : s == 1 ? off1 : off2) + i, src + i);
                    }
                }
            }
            asm volatile("cp.async.commit_group;" ::: "memory");
            return;
        }
        if (threadIdx.x == 0 && n_valid) {
            if (bulk_bytes) mbar_arrive_expect_tx(bar, bulk_bytes * n_valid); else mbar_arrive(bar);
        }
#pragma unroll
        for (int j = 0; j < R; ++j) {
            const long long rj = r + j * stride;
            if (rj >= n_rows) continue;
            float* dst0 = ring + (g * R + j) * buf_floats;
#pragma unroll
            for (int s = 0; s < 3; ++s) {
                const float* src = in.ptr[s] + rj * in.pitch[s];
                float* dst = dst0 + (s == 0 ? 0 : s == 1 ? off1 : off2);
                const int len = in.len[s];
                if (vec[s]) {
                    const int n4 = len >> 2;
                    if (threadIdx.x == 0 && n4)
                        bulk_g2s(static_cast<uint32_t>(__cvta_generic_to_shared(dst)), src, static_cast<uint32_t>(n4) * 16u, bar);
                    const int i = (n4 << 2) + static_cast<int>(threadIdx.x) - 32 * (s + 1);     // warp s + 1 fetches the tail
                    if (i >= (n4 << 2) && i < len) cp_async_4(dst + i, src + i);
                } else {
                    for (int i = threadIdx.x; i < len; i += PLIK_BIN_THREADS) cp_async_4(dst + i, src + i);
                }
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };

    // Tensor-core quadratic form: v_i = d_i sqrt(F_ii), scaled by a power of two q so that max |v_i| q lies in
    // [0.5, 1).  Written twice: fp32 (the epilogue's dot product) and as the contraction's A operand - split
    // fp16 hi/lo of 1024 q v in the swizzled K-major tile layout the fused kernel's loader streams.
    // Only the diff_pitch / 8 threads that own a 16-byte chunk of the operand tile take part: each reads its eight values
    // once and writes both forms (the fp32 row and the split tile) with 16-byte stores.
    // The R rows of a group are dealt to the R halves of the block (thread t: row t / 128, chunk t % 128), so that no warp
    // carries more than one row's write-out while the others wait at the barrier.
    static_assert(PLIK_BIN_THREADS % PLIK_R == 0, "rows of a group are dealt to equal parts of the block");
    constexpr int WT = PLIK_BIN_THREADS / PLIK_R;
    const int w_row = static_cast<int>(threadIdx.x) / WT, w_ch = static_cast<int>(threadIdx.x) % WT;
    const int tile_stage_off = (w_ch >> 2) * 16384, tile_kchunk = w_ch & 3;
    const long long tile_bytes = static_cast<long long>(diff_pitch / 32) * 16384;
    auto write_row = [&](long long rr, const float* vrow) {
        const float4 m0 = *reinterpret_cast<const float4*>(vrow + diff_pitch), m1 = *reinterpret_cast<const float4*>(vrow + diff_pitch + 4);
        const float m = fmaxf(fmaxf(fmaxf(m0.x, m0.y), fmaxf(m0.z, m0.w)), fmaxf(fmaxf(m1.x, m1.y), fmaxf(m1.z, m1.w)));
        // q = 2^-ex with m = f 2^ex, f in [0.5, 1): the exponent field alone (zero, subnormal, inf / NaN maxima keep q = 1)
        const uint32_t be = (__float_as_uint(m) >> 23) & 0xffu;
        const float q = (be >= 1u && be <= 253u) ? __uint_as_float((253u - be) << 23) : 1.f;
        const float4 a = *reinterpret_cast<const float4*>(vrow + 8 * w_ch), b = *reinterpret_cast<const float4*>(vrow + 8 * w_ch + 4);
        const float x[8] = {a.x * q, a.y * q, a.z * q, a.w * q, b.x * q, b.y * q, b.z * q, b.w * q};
        float* drow = diff + rr * diff_pitch + 8 * w_ch;
        *reinterpret_cast<float4*>(drow) = make_float4(x[0], x[1], x[2], x[3]);
        *reinterpret_cast<float4*>(drow + 4) = make_float4(x[4], x[5], x[6], x[7]);
        const int rt = static_cast<int>(rr & 127);
        uint4 hi, lo;
        plik_split2(x[0] * 1024.f, x[1] * 1024.f, hi.x, lo.x); plik_split2(x[2] * 1024.f, x[3] * 1024.f, hi.y, lo.y);
        plik_split2(x[4] * 1024.f, x[5] * 1024.f, hi.z, lo.z); plik_split2(x[6] * 1024.f, x[7] * 1024.f, hi.w, lo.w);
        uint8_t* t = tiles + (rr >> 7) * tile_bytes + tile_stage_off + rt * 64 + ((tile_kchunk ^ (rt >> 1)) & 3) * 16;
        *reinterpret_cast<uint4*>(t) = hi;
        *reinterpret_cast<uint4*>(t + 8192) = lo;
        if (w_ch == 0) row_scale[rr] = q;
    };
    // rows rr, rr + stride, ... binned in the previous iteration (their copies of the differences: parity `pv`)
    auto write_group = [&](long long rr, int pv) {
        if (w_ch >= diff_pitch / 8 || diff_pitch / 8 > WT) return;
        const long long rj = rr + w_row * stride;
        if (rj < n_rows) write_row(rj, vrow0 + (pv * R + w_row) * vstride);
    };

    long long r = blockIdx.x;
#pragma unroll
    for (int i = 0; i < G - 1; ++i) fetch(r + i * R * stride, i);
    uint32_t ring_s = static_cast<uint32_t>(__cvta_generic_to_shared(ring)), vrow0_s = static_cast<uint32_t>(__cvta_generic_to_shared(vrow0));
    uint32_t buf_bytes = 4u * buf_floats, vstride_bytes = 4u * vstride;
    asm volatile("" : "+r"(ring_s), "+r"(vrow0_s), "+r"(buf_bytes), "+r"(vstride_bytes));      // pinned in registers
#pragma unroll
    for (int p = 0; p < PLIK_MAX_PASSES; ++p) { my[p].off *= 4; asm volatile("" : "+r"(my[p].off), "+r"(my[p].bin)); }

    // one row: the thread's slots (loads of all passes first, then the arithmetic - the passes are independent chains; a pass
    // that does not exist has off = 0, weights 0 and bin = -1: it costs its instructions but keeps the code branch-free)
    auto bin_row = [&](long long rr, float c, uint32_t row_s, uint32_t vrow_s) {
        const float scale = units_factor / (c * c);
        float* brow = binned ? binned + rr * binned_pitch : nullptr;         // (optional) band powers of this row
        float* drow = TILES ? nullptr : diff + rr * diff_pitch;
        asm volatile("" : "+l"(brow), "+l"(drow));
        float v[PLIK_MAX_PASSES][PLIK_SEG];
#pragma unroll
        for (int p = 0; p < PLIK_MAX_PASSES; ++p) {
            const uint32_t qa = row_s + static_cast<uint32_t>(my[p].off);              // byte offset (scaled above)
#pragma unroll
            for (int j = 0; j < PLIK_SEG; ++j)                   // beyond the slot's end: neighbouring data or zero pad, weight 0
                asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v[p][j]) : "r"(qa + 4u * j));
        }
        float acc[PLIK_MAX_PASSES];
#pragma unroll
        for (int p = 0; p < PLIK_MAX_PASSES; ++p) {              // three partial sums of three: a 9-deep FMA chain is 36 cycles
            float a0 = v[p][0] * my[p].w[0], a1 = v[p][1] * my[p].w[1], a2 = v[p][2] * my[p].w[2];
#pragma unroll
            for (int j = 3; j < PLIK_SEG; j += 3) {
                a0 = fmaf(v[p][j], my[p].w[j], a0);
                if (j + 1 < PLIK_SEG) a1 = fmaf(v[p][j + 1], my[p].w[j + 1], a1);
                if (j + 2 < PLIK_SEG) a2 = fmaf(v[p][j + 2], my[p].w[j + 2], a2);
            }
            acc[p] = (a0 + a1) + a2;
        }
        float vmax = 0.f;
#pragma unroll
        for (int p = 0; p < PLIK_MAX_PASSES; ++p) {
            float sum = acc[p];
#pragma unroll
            for (int k = 1; k < PLIK_MAX_SLOTS_PER_BIN; ++k) {
                const float o = __shfl_down_sync(0xffffffffu, acc[p], k);
                if (k < my[p].n_slots) sum += o;
            }
            const bool owner = my[p].bin >= 0;
            const float model = sum * scale;
            if (brow && owner) brow[my[p].bin] = model;
            const float d = my[p].x_data - model;
            if (TILES) {
                const float vv = d * my[p].sqrt_f;
                if (owner) {
                    asm volatile("st.shared.f32 [%0], %1;" :: "r"(vrow_s + 4u * static_cast<uint32_t>(my[p].bin)), "f"(vv) : "memory");
                    vmax = fmaxf(vmax, fabsf(vv));
                }
            } else if (owner) drow[my[p].bin] = d;
        }
        if (TILES) {
            // warp maximum in one instruction: non-negative floats order like their bit patterns (a NaN sorts above inf and
            // leaves q = 1 in write_row, like inf)
            const uint32_t wm = __reduce_max_sync(0xffffffffu, __float_as_uint(vmax));
            if ((threadIdx.x & 31) == 0)
                asm volatile("st.shared.b32 [%0], %1;" :: "r"(vrow_s + 4u * static_cast<uint32_t>(diff_pitch) + 4u * (threadIdx.x >> 5)), "r"(wm) : "memory");
        }
    };

    int g = 0, par = 0;
    float c_next[R];
#pragma unroll
    for (int j = 0; j < R; ++j) c_next[j] = r + j * stride < n_rows ? cal[r + j * stride] : 1.f;
    long long r_prev = -1;                                       // first row of the group binned in the previous iteration, not yet written out
    uint32_t phases = 0;                                         // one parity bit per group
    for (; r < n_rows; r += R * stride) {
        asm volatile("cp.async.wait_group %0;" ::"n"(G - 2) : "memory");            // this thread's 4-byte copies of the group have landed
        while (!mbar_try_wait(bar0 + 8u * g, (phases >> g) & 1u)) { }              // ... and its bulk copies
        phases ^= 1u << g;
        float c[R];
#pragma unroll
        for (int j = 0; j < R; ++j) {                            // loaded one iteration ahead: a dependent global load per row
            c[j] = c_next[j];                                    // would sit on the critical path (9 % of the stall samples)
            const long long rn = r + (R + j) * stride;
            if (rn < n_rows) c_next[j] = cal[rn];
        }
        __syncthreads();                                         // the group is complete for every thread; every thread has binned the
                                                                 // previous group (its buffers are free, its differences and maxima are
                                                                 // complete) and written out the group before that (its copies are free)
        fetch(r + (G - 1) * R * stride, g == 0 ? G - 1 : g - 1);                   // refill the buffers of the previous group
#pragma unroll
        for (int j = 0; j < R; ++j)
            if (r + j * stride < n_rows)                          // uniform across the block
                bin_row(r + j * stride, c[j], ring_s + static_cast<uint32_t>(g * R + j) * buf_bytes,
                        vrow0_s + static_cast<uint32_t>(par * R + j) * vstride_bytes);
        if (TILES) {
            if (r_prev >= 0) write_group(r_prev, par ^ 1);
            r_prev = r;
            par ^= 1;
        }
        g = g + 1 == G ? 0 : g + 1;
    }
    if (TILES && r_prev >= 0) {
        __syncthreads();
        write_group(r_prev, par ^ 1);
    }
}

// partial[r][jt] = sum over the 128 columns n of block jt of d[r][n] * (sum_{k <= last row of block jt} d[r][k] F'[k][n])
// 128 x 128 output tile, 16-deep K steps double-buffered through registers, 8 x 8 accumulators per thread.
__global__ void __launch_bounds__(256)
plik_quadform_kernel(const float* __restrict__ D, int ldd, const float* __restrict__ Fp, int ldf, long long n_rows,
                     float* __restrict__ partial, int n_jt)
{
    __shared__ __align__(16) float As[2][PLIK_QF_BK][PLIK_QF_BM + 4];
    __shared__ __align__(16) float Bs[2][PLIK_QF_BK][PLIK_QF_BN + 4];
    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;            // 16 x 16 threads; rows ty*4.. and 64+ty*4.., columns tx*4.. and 64+tx*4..
    const int jt = blockIdx.x;
    const long long row0 = static_cast<long long>(blockIdx.y) * PLIK_QF_BM;
    const int col0 = jt * PLIK_QF_BN;
    const int k_end = col0 + PLIK_QF_BN;               // F' is upper triangular: rows k beyond this block's columns are zero
    const int n_kt = k_end / PLIK_QF_BK;

    float acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

    // global -> register staging: A tile 128 rows x 16 k (2 float4 per thread), B tile 16 k x 128 n (2 float4 per thread)
    float4 ra[2], rb[2];
    auto load_tiles = [&](int kt) {
        const int k0 = kt * PLIK_QF_BK;
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const int f = tid + u * 256, r = f >> 2, kq = (f & 3) * 4;
            const long long gr = row0 + r;
            ra[u] = gr < n_rows ? *reinterpret_cast<const float4*>(D + gr * ldd + k0 + kq) : make_float4(0.f, 0.f, 0.f, 0.f);
            const int kk = f >> 5, nq = (f & 31) * 4;
            rb[u] = *reinterpret_cast<const float4*>(Fp + static_cast<long long>(k0 + kk) * ldf + col0 + nq);
        }
    };
    auto store_tiles = [&](int buf) {
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const int f = tid + u * 256, r = f >> 2, kq = (f & 3) * 4;
            As[buf][kq][r] = ra[u].x; As[buf][kq + 1][r] = ra[u].y; As[buf][kq + 2][r] = ra[u].z; As[buf][kq + 3][r] = ra[u].w;
            const int kk = f >> 5, nq = (f & 31) * 4;
            *reinterpret_cast<float4*>(&Bs[buf][kk][nq]) = rb[u];
        }
    };
    load_tiles(0);
    store_tiles(0);
    __syncthreads();
    for (int kt = 0; kt < n_kt; ++kt) {
        const int buf = kt & 1;
        if (kt + 1 < n_kt) load_tiles(kt + 1);
#pragma unroll
        for (int kk = 0; kk < PLIK_QF_BK; ++kk) {
            const float4 a0 = *reinterpret_cast<const float4*>(&As[buf][kk][ty * 4]);
            const float4 a1 = *reinterpret_cast<const float4*>(&As[buf][kk][64 + ty * 4]);
            const float4 b0 = *reinterpret_cast<const float4*>(&Bs[buf][kk][tx * 4]);
            const float4 b1 = *reinterpret_cast<const float4*>(&Bs[buf][kk][64 + tx * 4]);
            const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
            const float w[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], w[j], acc[i][j]);
        }
        if (kt + 1 < n_kt) store_tiles(buf ^ 1);
        __syncthreads();
    }
    // epilogue: dot with this block's columns of d, reduce over the 16 threads that share a row
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const long long gr = row0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
        float s = 0.f;
        if (gr < n_rows) {
            const float4 d0 = *reinterpret_cast<const float4*>(D + gr * ldd + col0 + tx * 4);
            const float4 d1 = *reinterpret_cast<const float4*>(D + gr * ldd + col0 + 64 + tx * 4);
            s = acc[i][0] * d0.x + acc[i][1] * d0.y + acc[i][2] * d0.z + acc[i][3] * d0.w +
                acc[i][4] * d1.x + acc[i][5] * d1.y + acc[i][6] * d1.z + acc[i][7] * d1.w;
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


// tensor-core quadratic form: loglike = -0.5 chi^2 = -(acc_scale / q^2) * (sum of the row's partial dot products)
__global__ void plik_finish_tensor_kernel(const float* __restrict__ partial, int n_part, const float* __restrict__ row_scale,
                                          float acc_scale, long long n_rows, float* __restrict__ loglike)
{
    const long long r = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (r >= n_rows) return;
    float s = 0.f;
    for (int j = 0; j < n_part; ++j) s += partial[r * n_part + j];
    const float q = row_scale[r];
    loglike[r] = -(s * acc_scale) / (q * q);
}

}  // namespace cpb
