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

// Persistent blocks (three per SM) walking rows with a PLIK_RING-deep shared-memory ring: while a row is being binned
// the next two rows are already on their way (one bulk copy per input and row, completion on the buffer's mbarrier;
// each input starts 16-byte aligned in shared memory).  A slot is a run of at most PLIK_SEG consecutive multipoles of one bin; the host splits the 5 ... 33
// wide bins into slots so that the per-thread work is short, balanced and fully unrolled, and lays a bin's slots out on
// neighbouring lanes of one warp.  Every thread owns the same <= PLIK_MAX_PASSES slots for every row, so their window
// weights live in registers; the lane holding a bin's first slot adds the other partial sums (at most
// PLIK_MAX_SLOTS_PER_BIN - 1 shuffles, fixed order: deterministic) and writes d = X_data - model.
constexpr int PLIK_SEG = 9;
constexpr int PLIK_MAX_PASSES = 3;
constexpr int PLIK_MAX_SLOTS_PER_BIN = 4;
constexpr int PLIK_RING = 3;             // row buffers per block: one being binned, two in flight
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

// Shared memory of one block: PLIK_RING mbarriers (32 bytes), the row ring - each buffer holds the three inputs back to
// back (every one starting 16-byte aligned) plus PLIK_ROW_PAD zero floats, because a slot always reads PLIK_SEG
// elements (those beyond its end carry weight 0) -, the row's scaled differences and the per-warp maxima.
constexpr int PLIK_ROW_PAD = 8;
constexpr int PLIK_BAR_FLOATS = 8;

// TILES: tensor-core quadratic form behind it (scaled differences + operand tiles); false: plain differences for the fp32 one
template <bool TILES>
__global__ void __launch_bounds__(PLIK_BIN_THREADS, PLIK_BLOCKS_PER_SM)
plik_bin_kernel(PlikInputs in, const float* __restrict__ cal, long long n_rows,
                const PlikSlot* __restrict__ slots, int n_passes, float units_factor,
                float* __restrict__ diff, int diff_pitch, float* __restrict__ binned, long long binned_pitch,
                uint8_t* __restrict__ tiles, float* __restrict__ row_scale)
{
    extern __shared__ __align__(16) float plik_smem[];
    int off1 = (in.len[0] + 3) & ~3, off2 = off1 + ((in.len[1] + 3) & ~3);
    int buf_floats = off2 + ((in.len[2] + 3) & ~3) + PLIK_ROW_PAD;
    asm volatile("" : "+r"(off1), "+r"(off2), "+r"(buf_floats));      // pinned in registers: the compiler would otherwise re-derive
                                                                      // every loop invariant from the parameters in each iteration
    float* ring = plik_smem + PLIK_BAR_FLOATS;
    // tensor-core mode (tiles != nullptr): TWO copies of {the row's scaled differences, the per-warp maxima} - a row is
    // binned in one loop iteration and scaled / written out in the next, so one block barrier per row separates both the
    // ring hand-over and the row-maximum exchange (the first version needed two, and the kernel is bound by them)
    float* vrow0 = ring + PLIK_RING * buf_floats;
    const int vstride = diff_pitch + PLIK_BIN_THREADS / 32;
    const uint32_t bar0 = static_cast<uint32_t>(__cvta_generic_to_shared(plik_smem));
    // zero everything once: the pads between and after the inputs stay zero for the whole launch
    for (int i = threadIdx.x; i < PLIK_RING * buf_floats + 2 * vstride; i += PLIK_BIN_THREADS) ring[i] = 0.f;
    if (threadIdx.x == 0) {
        for (int i = 0; i < PLIK_RING; ++i) mbar_init(bar0 + 8u * i, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // the zero fill is ordered before the bulk copies' writes
    __syncthreads();

    // An input whose rows are 16-byte aligned arrives by one bulk copy per row (issued by thread 0, completion counted on
    // the buffer's mbarrier) plus at most three 4-byte cp.async for the elements beyond the last multiple of four; any
    // other input is gathered 4 bytes at a time by the whole block.
    bool vec[3];
    uint32_t bulk_bytes = 0;
#pragma unroll
    for (int s = 0; s < 3; ++s) {
        vec[s] = (in.pitch[s] & 3) == 0 && (reinterpret_cast<uintptr_t>(in.ptr[s]) & 15) == 0;
        if (vec[s]) bulk_bytes += static_cast<uint32_t>(in.len[s] >> 2) * 16u;
    }

    PlikSlot my[PLIK_MAX_PASSES];
#pragma unroll
    for (int p = 0; p < PLIK_MAX_PASSES; ++p) {
        if (p < n_passes) my[p] = slots[p * PLIK_BIN_THREADS + threadIdx.x];
        else { my[p].off = 0; my[p].last = 0; my[p].bin = -1; my[p].n_slots = 0; my[p].x_data = 0.f;
#pragma unroll
               for (int j = 0; j < PLIK_SEG; ++j) my[p].w[j] = 0.f; }
    }

    // (the kernel is bound by instruction issue - ncu, round 2: 65 % of the issue slots, 490 instructions per thread and row
    // in its first form - so everything a single warp can do is kept away from the other seven)
    int all_vec = vec[0] && vec[1] && vec[2];
    asm volatile("" : "+r"(all_vec), "+r"(bulk_bytes));                           // pinned: otherwise re-derived from the parameters in every iteration
    auto fetch = [&](long long r, int buf) {                    // always commits a group (empty past the last row)
        if (all_vec) {
            // warp 0 alone: lane 0 issues the three bulk copies, lanes 1 .. 9 the (at most three per input) tail elements
            if (threadIdx.x >= 32) return;
            if (r < n_rows) {
                float* dst0 = ring + buf * buf_floats;
                const uint32_t bar = bar0 + 8u * buf;
                if (threadIdx.x == 0) {
                    mbar_arrive_expect_tx(bar, bulk_bytes);
#pragma unroll
                    for (int s = 0; s < 3; ++s) {
                        const int n4 = in.len[s] >> 2;
                        if (n4) bulk_g2s(static_cast<uint32_t>(__cvta_generic_to_shared(dst0 + (s == 0 ? 0 : s == 1 ? off1 : off2))),
                                         in.ptr[s] + r * in.pitch[s], static_cast<uint32_t>(n4) * 16u, bar);
                    }
                } else if (threadIdx.x < 10) {
                    const int s = (static_cast<int>(threadIdx.x) - 1) / 3, u = (static_cast<int>(threadIdx.x) - 1) % 3;
                    const int len = s == 0 ? in.len[0] : s == 1 ? in.len[1] : in.len[2];
                    const int i = (len & ~3) + u;
                    if (i < len) {
                        const float* src = (s == 0 ? in.ptr[0] : s == 1 ? in.ptr[1] : in.ptr[2]) + r * (s == 0 ? in.pitch[0] : s == 1 ? in.pitch[1] : in.pitch[2]);
                        cp_async_4(dst0 + (s == 0 ? 0 : s == 1 ? off1 : off2) + i, src + i);
                    }
                }
            }
            asm volatile("cp.async.commit_group;" ::: "memory");
            return;
        }
        if (r < n_rows) {
            float* dst0 = ring + buf * buf_floats;
            const uint32_t bar = bar0 + 8u * buf;
            if (threadIdx.x == 0) {
                if (bulk_bytes) mbar_arrive_expect_tx(bar, bulk_bytes); else mbar_arrive(bar);
            }
#pragma unroll
            for (int s = 0; s < 3; ++s) {
                const float* src = in.ptr[s] + r * in.pitch[s];
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
    const int tile_stage_off = (static_cast<int>(threadIdx.x) >> 2) * 16384, tile_kchunk = static_cast<int>(threadIdx.x) & 3;
    const long long tile_bytes = static_cast<long long>(diff_pitch / 32) * 16384;
    auto write_row = [&](long long rr, const float* vrow) {
        if (static_cast<int>(threadIdx.x) >= diff_pitch / 8) return;
        const float4 m0 = *reinterpret_cast<const float4*>(vrow + diff_pitch), m1 = *reinterpret_cast<const float4*>(vrow + diff_pitch + 4);
        const float m = fmaxf(fmaxf(fmaxf(m0.x, m0.y), fmaxf(m0.z, m0.w)), fmaxf(fmaxf(m1.x, m1.y), fmaxf(m1.z, m1.w)));
        // q = 2^-ex with m = f 2^ex, f in [0.5, 1): the exponent field alone (zero, subnormal, inf / NaN maxima keep q = 1)
        const uint32_t be = (__float_as_uint(m) >> 23) & 0xffu;
        const float q = (be >= 1u && be <= 253u) ? __uint_as_float((253u - be) << 23) : 1.f;
        const float4 a = *reinterpret_cast<const float4*>(vrow + 8 * threadIdx.x), b = *reinterpret_cast<const float4*>(vrow + 8 * threadIdx.x + 4);
        const float x[8] = {a.x * q, a.y * q, a.z * q, a.w * q, b.x * q, b.y * q, b.z * q, b.w * q};
        float* drow = diff + rr * diff_pitch + 8 * threadIdx.x;
        *reinterpret_cast<float4*>(drow) = make_float4(x[0], x[1], x[2], x[3]);
        *reinterpret_cast<float4*>(drow + 4) = make_float4(x[4], x[5], x[6], x[7]);
        const int rt = static_cast<int>(rr & 127);
        uint4 hi, lo;
        plik_split2(x[0] * 1024.f, x[1] * 1024.f, hi.x, lo.x); plik_split2(x[2] * 1024.f, x[3] * 1024.f, hi.y, lo.y);
        plik_split2(x[4] * 1024.f, x[5] * 1024.f, hi.z, lo.z); plik_split2(x[6] * 1024.f, x[7] * 1024.f, hi.w, lo.w);
        uint8_t* t = tiles + (rr >> 7) * tile_bytes + tile_stage_off + rt * 64 + ((tile_kchunk ^ (rt >> 1)) & 3) * 16;
        *reinterpret_cast<uint4*>(t) = hi;
        *reinterpret_cast<uint4*>(t + 8192) = lo;
        if (threadIdx.x == 0) row_scale[rr] = q;
    };

    const long long stride = gridDim.x;
    long long r = blockIdx.x;
#pragma unroll
    for (int i = 0; i < PLIK_RING - 1; ++i) fetch(r + i * stride, i);
    uint32_t ring_s = static_cast<uint32_t>(__cvta_generic_to_shared(ring)), vrow0_s = static_cast<uint32_t>(__cvta_generic_to_shared(vrow0));
    uint32_t buf_bytes = 4u * buf_floats, vstride_bytes = 4u * vstride;
    asm volatile("" : "+r"(ring_s), "+r"(vrow0_s), "+r"(buf_bytes), "+r"(vstride_bytes));      // pinned in registers
#pragma unroll
    for (int p = 0; p < PLIK_MAX_PASSES; ++p) { my[p].off *= 4; asm volatile("" : "+r"(my[p].off), "+r"(my[p].bin)); }
    int buf = 0, par = 0;
    float c_next = r < n_rows ? cal[r] : 1.f;
    long long r_prev = -1;                                       // binned in the previous iteration, not yet written out
    uint32_t phases = 0;                                         // one parity bit per ring buffer
    for (; r < n_rows; r += stride) {
        asm volatile("cp.async.wait_group %0;" ::"n"(PLIK_RING - 2) : "memory");    // this thread's 4-byte copies of row r have landed
        while (!mbar_try_wait(bar0 + 8u * buf, (phases >> buf) & 1u)) { }          // ... and the row's bulk copies
        phases ^= 1u << buf;
        const float c = c_next;                                  // loaded one iteration ahead: a dependent global load per row
        if (r + stride < n_rows) c_next = cal[r + stride];       // would sit on the critical path (9 % of the stall samples)
        const float scale = units_factor / (c * c);
        float vmax = 0.f;
        __syncthreads();                                         // row r is complete for every thread; every thread has binned the
                                                                 // previous row (its ring buffer is free, its differences and maxima are
                                                                 // complete) and written out the row before that (its copy is free)
        fetch(r + (PLIK_RING - 1) * stride, buf == 0 ? PLIK_RING - 1 : buf - 1);     // refill the buffer of the previous row
        // shared-memory addresses as 32-bit integers behind explicit ld / st.shared: left to itself the compiler re-derives
        // `ring + buf * buf_floats + off` from the kernel parameters in every pass (a dozen instructions each time)
        const uint32_t row_s = ring_s + static_cast<uint32_t>(buf) * buf_bytes;
        const uint32_t vrow_s = vrow0_s + static_cast<uint32_t>(par) * vstride_bytes;
        float* vrow = vrow0 + par * vstride;
        float* brow = binned ? binned + r * binned_pitch : nullptr;          // (optional) band powers of this row
        float* drow = TILES ? nullptr : diff + r * diff_pitch;
        asm volatile("" : "+l"(brow), "+l"(drow));
        // all passes' loads first, then the arithmetic: the passes are independent chains, and a pass that does not exist
        // (p >= n_passes) has off = 0, weights 0 and bin = -1 - it costs its instructions but keeps the code branch-free
        float v[PLIK_MAX_PASSES][PLIK_SEG];
#pragma unroll
        for (int p = 0; p < PLIK_MAX_PASSES; ++p) {
            const uint32_t qa = row_s + static_cast<uint32_t>(my[p].off);              // byte offset (scaled before the loop)
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
            if ((threadIdx.x & 31) == 0) vrow[diff_pitch + (threadIdx.x >> 5)] = __uint_as_float(wm);
            if (r_prev >= 0) write_row(r_prev, vrow0 + (par ^ 1) * vstride);
            r_prev = r;
            par ^= 1;
        }
        buf = buf + 1 == PLIK_RING ? 0 : buf + 1;
    }
    if (TILES && r_prev >= 0) {
        __syncthreads();
        write_row(r_prev, vrow0 + (par ^ 1) * vstride);
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
