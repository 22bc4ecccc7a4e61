// Persistent fused-MLP kernel for sm_100a (CPB200_PREC_F16X3).
//
// 74 CTA pairs (cluster of 2, one CTA per SM) walk pairs of 128-row tiles through the WHOLE network
// (standardisation, every dense layer, activation, PCA rescale + basis contraction,
// de-standardisation, 10**x).  Every dense contraction runs on the 5th-gen tensor cores
// (tcgen05.mma.cta_group::2 kind::f16, M = 256 across the pair, N <= 256, fp32 accumulators in each
// CTA's TMEM) as a 3-product fp16 split:
//     A.W  ~=  A_hi.W_hi + A_lo.W_hi + A_hi.W_lo      (x_hi = fp16(x), x_lo = fp16(x - x_hi); the A tiles hold -A_lo and the
//                                                      middle product is issued with the negate-A descriptor bit)
// which keeps ~22 significand bits per operand (single-pass bf16/tf32 miss the 1e-5 tolerance,
// SURVEY 7.2).
//
// Data movement: operands are stored in global memory already in the tensor core's shared-memory
// layouts - weights as K-major no-swizzle core matrices (8 rows x 16 bytes), activations as K-major
// SWIZZLE_64B tiles (64-byte rows) - so a pipeline stage is two plain bulk copies (cp.async.bulk,
// the TMA engine's 1-D mode, no tensor maps): 16 KB of activations (hi|lo, 128 rows x 32 k) and this
// CTA's half of the weight stage (NC*64 B).  Activations between layers go through a CTA-private
// scratch that lives in L2 (a 128x512 tile in split precision is 256 KB: more than shared memory, and
// TMEM holds the accumulators); the epilogue writes it with full-line 16-byte stores and the loader
// streams it back.  Two row tiles are interleaved per CTA so the epilogue of one overlaps the MMAs of
// the other; the order of jobs is a host-built table (layer 0 of the next pair is merged into the
// current pair's output phase).
//
// Warp roles (384 threads): warp 0 loader (bulk copies), warp 1 MMA issuer (leader CTA; stage relay in
// the peer) - both walk their loops as whole warps with warp-uniform values and let one elected lane
// issue, so addresses and descriptors stay in uniform registers -, warp 2 TMEM allocator + parameter
// standardiser, warp 3 idle, warps 4..11 epilogue (2 warps per TMEM lane quarter, 32-column pieces
// dealt alternately, TMEM reads software-pipelined half a piece ahead).
// The epilogue reads accumulators with tcgen05.ld.16x256b, whose fragment gives a thread 2 adjacent
// columns of 2 rows; the weight columns are packed in that order, so 4 neighbouring lanes hold one
// 64-byte tile row and the next layer's operand is written with full 128-byte lines - no
// shared-memory transpose (the shared-memory data pipe is the scarce resource: the tensor core's
// operand reads and the bulk copies already use most of it).
//
// The same pipeline also evaluates the Planck-lite chi^2 (EPI_QUADFORM): one upper-triangular contraction
// whose layer-0 operand tiles are prebuilt by plik_bin_kernel (UmmaParams::ext_a0), see plik_kernels.cuh.
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
constexpr int UM_MAX_NC = 256;                               // widest N chunk (one accumulator = two TMEM slots)
constexpr int UM_SLOT_COLS = 128;                            // TMEM is handed out in four slots of 128 columns
constexpr int UM_TMEM_SLOTS = 4;
constexpr int UM_STAGES = 6;
constexpr int UM_ACT_BUFS = 3;                               // activation buffers per CTA (rotated between its two tiles)
constexpr int UM_CLUSTER = 2;                                // CTA pair: one tcgen05.mma.cta_group::2 (M = 256) drives both SMs
constexpr int UM_A_HALF_BYTES = UM_TILE_M * UM_KS * 2;       // 8192: one of (hi, lo)
constexpr int UM_A_STAGE_BYTES = 2 * UM_A_HALF_BYTES;        // 16384
constexpr int UM_B_STAGE_BYTES_MAX = 2 * (UM_MAX_NC / 2) * UM_KS * 2;   // 16384: each CTA of the pair holds half of the chunk's columns
constexpr int UM_STAGE_BYTES = UM_A_STAGE_BYTES + UM_B_STAGE_BYTES_MAX;   // 32768
constexpr int UM_MAX_GEMMS = 16;                             // contractions of all models of one launch (job-table field: 4 bits)
constexpr int UM_MAX_MODELS = 3;                             // emulators evaluated on the same rows by one launch (TT + TE + EE)
constexpr int UM_THREADS = 384;
constexpr int UM_EPI_WARP0 = 4;
constexpr int UM_EPI_WARPS = 8;
constexpr int UM_EPI_THREADS = UM_EPI_WARPS * 32;
constexpr bool UM_DISCARD_DEAD_TILES = true;                      // discard.global.L2 on consumed activation tiles
constexpr int UM_PIECE = 32;                                 // accumulator columns per epilogue piece
constexpr float UM_INPUT_SCALE = 1024.f;                     // power-of-two pre-scale of the standardised parameters
constexpr float UM_TRUNC_BIAS_ULPS_PER_STEP = 0.16f;               // mean accumulator loss per MMA step, in ulps (calibrated on B200)
constexpr float UM_ACT_SCALE = 64.f;                         // power-of-two pre-scale of every produced activation

// shared memory carve-up (bytes from the 1024-aligned base)
constexpr int UM_SMEM_STAGES = 0;
#ifndef UM_PAR_SLOTS_N
#define UM_PAR_SLOTS_N 6
#endif
constexpr int UM_PAR_SLOTS = UM_PAR_SLOTS_N;          // ring of per-chunk epilogue constants: deep enough that the loader never waits for an epilogue
constexpr int UM_SMEM_EPIPAR = UM_STAGES * UM_STAGE_BYTES;                   // UM_PAR_SLOTS x 256 float4
constexpr int UM_SMEM_BARS = UM_SMEM_EPIPAR + UM_PAR_SLOTS * UM_MAX_NC * 16;
constexpr int UM_NUM_BARS = 2 * UM_STAGES + 2 * UM_TMEM_SLOTS + 4 + 4 * UM_MAX_MODELS + 2 * UM_PAR_SLOTS;
constexpr int UM_SMEM_TMEMPTR = UM_SMEM_BARS + UM_NUM_BARS * 8;
constexpr int UM_MAX_JOBS = 320;         // entries of the job table (parts of jobs) per iteration
constexpr int UM_SMEM_BYTES = UM_SMEM_TMEMPTR + 16 + 1024;                   // + alignment slack

struct UmmaGemm {
    const __half* wpack;   // [n_chunks][k_stages][CTA0: hi|lo half tiles][CTA1: hi|lo half tiles], core-matrix order
    const float4* epi_b;   // last contraction of an emulator with band-power binning attached (cpb200_engine_set_binning): per
                           // column {scale * log2 10, shift * log2 10, window weight if the column falls into the FIRST bin its
                           // group of four columns touches, window weight if into the SECOND}, lane-interleaved like ACT
    const float4* epi;     // per padded output column (lane-interleaved inside 32-column blocks):
                           // ACT {b, -alpha*log2e, beta*s, (1-beta)*s}; AFFINE {scale, shift}; OUT {scale, shift, *log2 10}
    int k_stages;          // K_pad / 32
    int nc;                // chunk width: multiple of 32, <= 256
    int n_chunks;
    int epi_kind;          // cpb::Epilogue
    float acc_scale;       // undoes the power-of-two operand pre-scales on the accumulator (EPI_ACT)
    int tri;               // upper-triangular weights: chunk c only needs the K stages up to its last column
    short model;           // which emulator of the launch this contraction belongs to
    short layer0;          // first contraction of its emulator: its operand comes from the standardiser
};

// one emulator of a launch: its parameter columns, standardisation constants and where its spectra go
struct UmmaModel {
    const float* params;       // (n_rows, n_parameters)
    const float* pmean;
    const float* pstd;
    float* out;
    long long out_pitch;
    int n_parameters;
    int n_modes;
    int ten_to;
    int accumulate;            // output layers add the log10 values already in `out` before the optional 10**
    int binned;                // output layers write per-4-column partial band powers of 10**y instead of the spectrum:
                               // row r, columns 4g .. 4g+3 -> out[r * pitch + 2g + {0, 1}] (first / second bin of the group)
};

struct UmmaParams {
    UmmaGemm g[UM_MAX_GEMMS];
    int n_gemms;
    int n_models;
    UmmaModel m[UM_MAX_MODELS];
    int discard_ok;            // every operand-producing contraction has a chunk width that is a multiple of 64 (see the epilogue's discard)
    int debug_flags;           // profiling instantiation only: 1 = epilogues skip all work, 2 = epilogues only read TMEM, 64 = no MMAs, 128 = no operand copies, 256 = no discards, 512 = the OUTPUT epilogues store nothing (results are garbage)
    long long n_rows;
    long long n_pairs;         // ceil(n_tiles / 2)
    long long n_pairgroups;    // ceil(n_pairs / UM_CLUSTER): one pair per CTA of a cluster, processed in lockstep
    int out_split;             // small batches: this many clusters share a pair-group - each walks the whole trunk (its own
                               // copy of the activations) but only every out_split-th chunk of the output contraction, so the
                               // spectrum's columns are produced by otherwise idle SMs (1 = off; grid = clusters x out_split x 2)
    uint8_t* scratch;          // gridDim.x * scratch_cta_bytes
    unsigned long long scratch_cta_bytes;
    unsigned a0_bytes;         // per tile: k_stages(gemm 0) * 16384
    unsigned abuf_bytes;       // per activation buffer: max k_stages(gemm >= 1) * 16384; a CTA owns UM_ACT_BUFS of them
    // Job table (cpb200_api.cu build_jobs): one entry per PART of a job.  A job = (contraction g, tile t, N-chunk c)
    // = one accumulator; its K loop may be cut into parts so that another job's MMAs are issued in between (the static
    // schedule interleaves short jobs into long ones).  Lives in the constant bank with the
    // other parameters, so every role decodes it with uniform loads.  Encoding: um_job_w0 / um_job_w1 below.
    unsigned jobs[UM_MAX_JOBS];
    unsigned short jobs_x[UM_MAX_JOBS];
    int n_jobs;
    int rot_mul;               // rotating activation buffers: buffer = (static index + pair number * rot_mul) mod UM_ACT_BUFS
    unsigned pbuf_bytes;       // per-pair-parity operand buffers of the last contraction (overlapped output phase); 0 = none
    int* error_flag;           // host-mapped status words: [0] set to a role/barrier code by the watchdog before trapping,
                               // [1] set to 1 when a split operand left the fp16 range (the batch's result is invalid)
    unsigned long long* max_cycles;
    // quadratic-form mode (EPI_QUADFORM, single contraction): the layer-0 operand tiles are prebuilt in global memory
    // (tile i at ext_a0 + i * a0_bytes, the layout the standardiser would write), qf_d holds the same rows in fp32
    // (row pitch qf_pitch, qf_cols valid columns) and qf_partial receives 2 * n_chunks partial sums per row
    const uint8_t* ext_a0;
    const float* qf_d;
    long long qf_pitch;
    int qf_cols;
    float* qf_partial;   // every launch: max over CTAs of the SM cycles between kernel entry and exit
    long long* counters;       // optional debug counters, 16 per CTA (nullptr = off); see UM_CNT_*
    // profiling probe (CPB200_ALIGN_EVERY, PROF instantiation only): every `align_every` pairs the MMA issuers of all clusters
    // meet at a counter in global memory, so that the clusters walk their trunks and their output phases at the same time
    int align_every;
    unsigned* align_counter;
};

// job-table entry, word 0: which part of which job, and where its accumulator lives
//   g (4 bits) | t << 4 | c << 5 (7 bits) | nxt << 12 | s0 << 13 (6) | ns << 19 (6) | first << 25 | last << 26 | slot << 27 (2) | two << 29
// nxt: the entry belongs to the pair of iteration `it` (else to pair it-1); stages [s0, s0 + ns) of the job's K loop;
// first / last part of the job; TMEM slot (128 columns) of the accumulator, `two` = it spans slot and slot + 1.
__host__ __device__ inline unsigned um_job_w0(int g, int t, int c, int nxt, int s0, int ns, int first, int last, int slot, int two) {
    return static_cast<unsigned>(g | (t << 4) | (c << 5) | (nxt << 12) | (s0 << 13) | (ns << 19) | (first << 25) | (last << 26) | (slot << 27) | (two << 29));
}
// word 1: operand buffer the job reads | buffer its epilogue writes << 8.  Codes: UM_BUF_A0 + t = layer-0 operand of tile t;
// UM_BUF_ROT + i = rotating activation buffer (i + pair * rot_mul) mod 3; UM_BUF_PAR + t = buffer t of the pair's parity set
enum { UM_BUF_NONE = 0, UM_BUF_A0 = 0x10, UM_BUF_ROT = 0x20, UM_BUF_PAR = 0x40 };
__host__ __device__ inline unsigned short um_job_w1(int in_code, int out_code) { return static_cast<unsigned short>(in_code | (out_code << 8)); }

// debug counter slots (cycles), per CTA
enum { UM_CNT_LOAD_WAIT_EMPTY = 0, UM_CNT_LOAD_WAIT_AREADY = 1, UM_CNT_LOAD_TOTAL = 2,
       UM_CNT_MMA_WAIT_FULL = 3, UM_CNT_MMA_WAIT_ACCEMPTY = 4, UM_CNT_MMA_TOTAL = 5,
       UM_CNT_EPI_WAIT_ACCFULL = 6, UM_CNT_EPI_ACT = 7, UM_CNT_EPI_OUT = 8, UM_CNT_EPI_TOTAL = 9,
       UM_CNT_IN_WAIT = 10, UM_CNT_IN_TOTAL = 11, UM_CNT_SLOTS = 16 };
// After the per-CTA slots the counter buffer holds UM_MAX_JOBS rows of UM_CNT_SLOTS values written by CTA 0
// only: per job-table entry, waits summed over all iterations plus one iteration's time stamps.
enum { UM_JOB_MMA_WAIT_FULL = 0, UM_JOB_MMA_WAIT_ACCEMPTY = 1, UM_JOB_MMA_BUSY = 2, UM_JOB_EPI_WAIT = 3,
       UM_JOB_EPI_BUSY = 4, UM_JOB_LOAD_WAIT_AREADY = 5, UM_JOB_LOAD_WAIT_EMPTY = 6,
       UM_JOB_T_MMA0 = 8, UM_JOB_T_MMA1 = 9, UM_JOB_T_EPI0 = 10, UM_JOB_T_EPI1 = 11, UM_JOB_T_LOAD0 = 12,
       UM_JOB_T_LOAD1 = 13, UM_JOB_TRACE_ITERATION = 3, UM_TRACE_STAGES = 16 };
// ... followed by UM_MAX_JOBS x UM_TRACE_STAGES x 2 time stamps of the traced iteration: per K stage, when the
// loader issued its copies and when the MMA issuer saw them landed.

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
// Pure event signal: no data written by this thread is published through the barrier.  The default form has release
// semantics and compiles to MEMBAR.ALL.CTA + SYNCS.ARRIVE: the fence waits until every store the thread has in flight is
// performed - at the end of an epilogue job that is the whole burst of spectrum / activation stores still queued in the
// LSU (ncu source view, round 2: 9 % of the cmb_PP launch's warp samples sat on that MEMBAR), and the warp could not
// start the next job's TMEM loads meanwhile.  Handing back TMEM (its reads completed with tcgen05.wait::ld) or a
// constants slot (its LDS results are already consumed) needs no such ordering.
__device__ __forceinline__ void mbar_arrive_relaxed(uint32_t bar) {
    asm volatile("mbarrier.arrive.relaxed.cta.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
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
    const long long t0 = clock64();           // try_wait itself may suspend the thread: time the whole wait
    if (mbar_try_wait(bar, parity)) return clock64() - t0;
    mbar_wait_slow(bar, parity, error_flag, code);
    return clock64() - t0;
}
// true in exactly one lane of a converged warp; code under it can use the warp's uniform values directly
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n.reg .pred p;\nelect.sync _|p, 0xffffffff;\nselp.u32 %0, 1, 0, p;\n}" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst_smem, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst_smem), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
// L2 eviction-priority policies: the CTA-private activation scratch and the weights must stay in L2
// (evict_last); an activation tile's final read and the spectrum output are streaming (evict_first)
__device__ __forceinline__ uint64_t policy_evict_last() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ uint64_t policy_evict_first() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ void bulk_g2s_hint(uint32_t dst_smem, const void* src, uint32_t bytes, uint32_t bar, uint64_t pol) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
                 ::"r"(dst_smem), "l"(src), "r"(bytes), "r"(bar), "l"(pol) : "memory");
}
__device__ __forceinline__ void st_global_v4_hint(void* ptr, const uint4& v, uint64_t pol) {
    asm volatile("st.global.L2::cache_hint.v4.b32 [%0], {%1, %2, %3, %4}, %5;"
                 ::"l"(ptr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w), "l"(pol) : "memory");
}
// drop a dead 128-byte line from L2 without writing it back to HBM
__device__ __forceinline__ void l2_discard_line(const void* ptr) {
    asm volatile("discard.global.L2 [%0], 128;" ::"l"(ptr) : "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async;" ::: "memory"); }
// global-only form: orders this thread's generic-proxy global stores before async-proxy (bulk copy) reads of them.
// SASS: a single FENCE.VIEW.ASYNC.G, where the unqualified form is MEMBAR.ALL.GPU + FENCE.VIEW.ASYNC.S.
__device__ __forceinline__ void fence_proxy_async_global() {
    asm volatile("fence.proxy.async.global;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// 2-CTA MMA: D[256 x N] (128 rows in each CTA's TMEM) += A[256 x 16] (128 rows from each CTA's smem)
// x B[N x 16] (N/2 rows from each CTA's smem).  Issued by one thread of the leader CTA only.
__device__ __forceinline__ void umma_f16_2cta(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                              uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
// arrive, once all MMAs issued so far have completed, on the barrier at this offset in both CTAs of the pair
__device__ __forceinline__ void umma_commit_2cta(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(bar), "h"(static_cast<uint16_t>(3)) : "memory");
}
// arrive on the barrier at the same shared-memory offset in another CTA of the cluster
__device__ __forceinline__ uint32_t map_to_cta(uint32_t saddr, uint32_t target_cta) {
    uint32_t raddr;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(raddr) : "r"(saddr), "r"(target_cta));
    return raddr;
}
// pure event signal (no data of this thread is published through it): no release fence on the issue path
__device__ __forceinline__ void mbar_arrive_remote_relaxed(uint32_t raddr) {
    asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(raddr) : "memory");
}
// 16 lanes x (4 x 256 bit): thread t gets, for column group i in 0..3: v[4i+0..1] = (lane t/4, cols 8i+2(t%4)+{0,1}),
// v[4i+2..3] = (lane t/4+8, same cols)
__device__ __forceinline__ void tmem_ld_16x256b_x4(uint32_t taddr, uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.16x256b.x4.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr) : "memory");
}
// wait for outstanding tcgen05.ld; the registers are listed as in/out so that no use of them can be
// scheduled above the wait
__device__ __forceinline__ void tmem_ld_wait16(uint32_t (&v)[16]) {
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(v[0]), "+r"(v[1]), "+r"(v[2]), "+r"(v[3]), "+r"(v[4]), "+r"(v[5]), "+r"(v[6]), "+r"(v[7]),
                   "+r"(v[8]), "+r"(v[9]), "+r"(v[10]), "+r"(v[11]), "+r"(v[12]), "+r"(v[13]), "+r"(v[14]), "+r"(v[15])
                 :: "memory");
}
// Ordering of the epilogue's plain stores to the L2-resident scratch before the loader's bulk
// (async-proxy) reads of the same lines by the same CTA.
#ifndef UM_SCRATCH_FENCE
#define UM_SCRATCH_FENCE() ((void)0)   /* fence.proxy.async alone orders them; validated by the 1M-row cross-check tests */
#endif

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
// Activation (A) tiles use the K-major SWIZZLE_64B layout instead: a stage is 128 rows x 32 k = 64-byte
// rows, 8-row atoms of 512 B (SBO), 16-byte chunks XOR-ed with (row>>1)&3.  Rows are contiguous, so the
// epilogue's fragment (4 lanes = one row's 64 B, 8 lanes = two rows) stores whole 128-byte lines.
__device__ __forceinline__ uint64_t umma_desc_a(uint32_t saddr) {
    uint64_t d = 0;
    d |= static_cast<uint64_t>((saddr & 0x3FFFFu) >> 4);
    d |= static_cast<uint64_t>(1) << 16;               // LBO unused for swizzled K-major
    d |= static_cast<uint64_t>(512u >> 4) << 32;       // SBO: 8 rows x 64 B
    d |= static_cast<uint64_t>(1) << 46;
    d |= static_cast<uint64_t>(4) << 61;               // SWIZZLE_64B
    return d;
}
// byte offset of the 16-byte chunk holding k..k+7 (k multiple of 8, < 32) of row r inside one A tile
__device__ __forceinline__ uint32_t a_tile_off(int r, int kchunk) {
    return static_cast<uint32_t>(r) * 64u + static_cast<uint32_t>((kchunk ^ (r >> 1)) & 3) * 16u;
}
// kind::f16 instruction descriptor: D=f32, A=B=f16, both K-major, M=256 across the CTA pair, N=nc
__device__ __forceinline__ uint32_t umma_idesc(int nc) {
    return (1u << 4) | (static_cast<uint32_t>(nc >> 3) << 17) | (static_cast<uint32_t>((UM_CLUSTER * UM_TILE_M) >> 4) << 24);
}
constexpr uint32_t UM_IDESC_NEG_A = 1u << 13;     // negate the A operand: the lo halves of the A tiles are stored negated (split2n)

// split two floats into fp16 hi and lo parts (hi = rn(x), lo = rn(x - hi)), packed as half2.  The conversions
// saturate (F2FP.SATFINITE, same cost as the plain form): a value beyond the fp16 range becomes +-65504 instead of
// hi = inf, lo = -inf, which would turn the whole row into NaN one contraction later.  Such a row is still wrong -
// the producers below flag it (UmmaParams::error_flag[1]) and the host reports CPB200_E_RANGE.
constexpr float UM_F16_MAX = 65504.f;
__device__ __forceinline__ uint32_t pack_f16x2_sat(float a, float b) {
    uint32_t h;
    asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(h) : "f"(b), "f"(a));   // first source -> upper half
    return h;
}
// Packed fp32 pairs (sm_100: FFMA2 / FMUL2 / FADD2 work on an aligned 64-bit register pair): the epilogues are bound by
// instruction issue next to the special-function unit, and a pair instruction does two lanes' worth of work per slot.
__device__ __forceinline__ uint64_t pk2(float x, float y) {
    uint64_t r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(x), "f"(y));
    return r;
}
__device__ __forceinline__ void upk2(uint64_t v, float& x, float& y) {
    asm("mov.b64 {%0, %1}, %2;" : "=f"(x), "=f"(y) : "l"(v));
}
__device__ __forceinline__ uint64_t ffma2(uint64_t a, uint64_t b, uint64_t c) {
    uint64_t d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
__device__ __forceinline__ uint64_t fmul2(uint64_t a, uint64_t b) {
    uint64_t d;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ uint64_t fadd2(uint64_t a, uint64_t b) {
    uint64_t d;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
// Split two floats into fp16 hi and the NEGATED lo part, packed as half2 pairs: hi = rn(x), nlo = rn(hi - x).
// `sub.f32.f16` (sm_100 mixed-precision add: SASS FHADD) widens the half operand inside the instruction, so the
// residual costs one instruction per value instead of an unpack and a subtract; the difference hi - x is exact in
// fp32.  The tensor core takes the sign back: the A_lo product is issued with the descriptor's negate-A bit
// (umma_idesc | UM_IDESC_NEG_A), so every A operand tile in scratch holds [hi | -lo].
__device__ __forceinline__ void split2n(float a, float b, uint32_t& hi, uint32_t& nlo) {
    asm("{\n\t.reg .b16 l, u;\n\t.reg .f32 n0, n1;\n\t"
        "cvt.rn.satfinite.f16x2.f32 %0, %3, %2;\n\t"      // first source -> upper half
        "mov.b32 {l, u}, %0;\n\t"
        "sub.rn.f32.f16 n0, l, %2;\n\t"
        "sub.rn.f32.f16 n1, u, %3;\n\t"
        "cvt.rn.satfinite.f16x2.f32 %1, n1, n0;\n\t}"
        : "=&r"(hi), "=r"(nlo) : "f"(a), "f"(b));
}
__device__ __forceinline__ void split2n(uint64_t ab, uint32_t& hi, uint32_t& nlo) {
    float a, b;
    upk2(ab, a, b);
    split2n(a, b, hi, nlo);
}
#ifndef UM_EARLY_RELEASE
#define UM_EARLY_RELEASE 0    // 1: a two-slot accumulator's first slot is handed back as soon as its columns are drained (measured
                              // on B200: the mid-job fence costs cmb_PP 3.5 % and no schedule in use needs the half accumulator)
#endif
#ifndef UM_ACC_HOIST
#define UM_ACC_HOIST 0
#endif
#ifndef UM_EPI_INLINE
#define UM_EPI_INLINE __forceinline__
#endif
#ifndef UM_RANGE_CHECK
#define UM_RANGE_CHECK 1      // producers of split operands report values beyond the fp16 range (0: saturate silently)
#endif
#ifndef UM_ACCURATE_ACT
#define UM_ACCURATE_ACT 0     // 1 (accuracy study only): exp2f / IEEE division instead of the MUFU approximations
#endif
__device__ __forceinline__ float ex2_approx(float x) {
    if (UM_ACCURATE_ACT) return exp2f(x);
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float rcp_approx(float x) {
    if (UM_ACCURATE_ACT) return __frcp_rn(x);
    float y;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
// ---------------------------------------------------------------------------------------------
// epilogue of one accumulator chunk for one warp (templated on the layer kind so that the inner
// loops carry exactly one code path)
// ---------------------------------------------------------------------------------------------
struct EpiCtx {
    uint32_t t_addr0;        // TMEM address of this warp's first column, lanes +0
    int nc;                  // chunk width (multiple of 32)
    int cw;                  // this warp takes the 32-column pieces cw, cw + UM_EPI_WARPS/4, ...
    int chunk_col0;          // c * nc
    int q, tq, tr;
    const float4* par;       // the chunk's per-column constants in shared memory (logical column order)
    int dbg;                 // profiling ablations (always 0 in the production instantiation): 4 = no global stores, 8 = no constant loads
    // TMEM hand-back: the accumulator sits in slot `slot` (and slot + 1 when `two`); the leader CTA's MMA issuer owns
    // the "slot empty" barriers, the peer arrives on them remotely
    int lane;
    uint32_t rel_bar;        // two-slot accumulator: the leader's "slot empty" barrier of its FIRST slot (local address in the
                             // leader, cluster-mapped in the peer), handed back as soon as its columns are drained; 0 = none
    bool leader;
};

// all lanes of the warp have finished reading TMEM slot `sl` (tcgen05.wait::ld precedes this): give it back
__device__ __forceinline__ void epi_release_first_slot(const EpiCtx& E) {
    tc_fence_before();
    __syncwarp();
    if (E.lane == 0) {
        if (E.leader) mbar_arrive_relaxed(E.rel_bar);
        else mbar_arrive_remote_relaxed(E.rel_bar);        // pure event: the TMEM reads are complete
    }
}
// A two-slot accumulator's first slot is drained once a warp's next piece lies in the second slot: the piece loops below
// run in two legs (columns below / from UM_SLOT_COLS) and hand the first slot back in between, so that a short job can
// be issued into it while this epilogue continues.

// hidden layers (KIND = EPI_ACT) and the PCA-coefficient rescale (EPI_AFFINE): produce the next
// contraction's A operand, split into fp16 hi / -lo, in the swizzled K-major tile layout.
// One half piece = 16 TMEM lanes (rows tq, tq+8 of this warp's lower or upper 16 rows) x 32 accumulator columns:
// (v[4i + 2rr], v[4i + 2rr + 1]) = row 8rr + tq, the lane's logical columns 2i, 2i + 1 - an aligned register pair, so
// the whole chain runs on packed fp32 pairs.  Constants per column pair i: pa[i] = {b0, b1, a0, a1} (bias,
// -alpha*log2e), pb[i] = {c0, c1, d0, d1} (beta*s, (1-beta)*s); EPI_AFFINE: pa[i] = {scale0, scale1, shift0, shift1}.
// All eight pairs of the half piece walk the activation stage by stage together: 16 independent chains, so the
// special-function unit always has a queued instruction while the FMA pipe works on the other stages.
template <int KIND>
__device__ __forceinline__ void epi_operand_half(const uint32_t (&v)[16], const float4 (&pa)[4], const float4 (&pb)[4], uint64_t scale2,
                                                 uint8_t* dst, uint64_t pol_keep, int dbg, int* status) {
    uint64_t h[8];                                  // h[2i + rr]
    if (KIND == EPI_ACT) {
        uint64_t z[8];
        float e[16];
#pragma unroll
        for (int p = 0; p < 8; ++p) {
            const int i = p >> 1, rr = p & 1;
            z[p] = ffma2(pk2(__uint_as_float(v[4 * i + 2 * rr]), __uint_as_float(v[4 * i + 2 * rr + 1])), scale2, pk2(pa[i].x, pa[i].y));
        }
#pragma unroll
        for (int p = 0; p < 8; ++p) upk2(fmul2(z[p], pk2(pa[p >> 1].z, pa[p >> 1].w)), e[2 * p], e[2 * p + 1]);
#pragma unroll
        for (int k = 0; k < 16; ++k) e[k] = ex2_approx(e[k]);
        const uint64_t one2 = pk2(1.f, 1.f);
#pragma unroll
        for (int p = 0; p < 8; ++p) upk2(fadd2(pk2(e[2 * p], e[2 * p + 1]), one2), e[2 * p], e[2 * p + 1]);
#pragma unroll
        for (int k = 0; k < 16; ++k) e[k] = rcp_approx(e[k]);
#pragma unroll
        for (int p = 0; p < 8; ++p) {
            const int i = p >> 1;
            h[p] = fmul2(ffma2(pk2(pb[i].z, pb[i].w), pk2(e[2 * p], e[2 * p + 1]), pk2(pb[i].x, pb[i].y)), z[p]);
        }
    } else {
#pragma unroll
        for (int p = 0; p < 8; ++p) {
            const int i = p >> 1, rr = p & 1;
            h[p] = ffma2(pk2(__uint_as_float(v[4 * i + 2 * rr]), __uint_as_float(v[4 * i + 2 * rr + 1])), pk2(pa[i].x, pa[i].y), pk2(pa[i].z, pa[i].w));
        }
    }
    if (UM_RANGE_CHECK) {
        // one FMNMX3 per pair on the otherwise idle ALU pipe; NaN cannot occur (exp overflow gives sigma = 0)
        float m = 0.f;
#pragma unroll
        for (int p = 0; p < 8; ++p) { float x, y; upk2(h[p], x, y); m = fmaxf(m, fmaxf(fabsf(x), fabsf(y))); }
        if (m >= UM_F16_MAX && status) *reinterpret_cast<volatile int*>(status + 1) = 1;   // same value from every writer: no atomic
    }
#pragma unroll
    for (int rr = 0; rr < 2; ++rr) {
        uint4 hi, lo;
        split2n(h[0 + rr], hi.x, lo.x);
        split2n(h[2 + rr], hi.y, lo.y);
        split2n(h[4 + rr], hi.z, lo.z);
        split2n(h[6 + rr], hi.w, lo.w);
        if (dbg & 4) { if (hi.x == 0x7fff7fffu && lo.y == 0x12345u) st_global_v4_hint(dst, hi, pol_keep); continue; }
        st_global_v4_hint(dst + rr * 512, hi, pol_keep);
        st_global_v4_hint(dst + rr * 512 + UM_A_HALF_BYTES, lo, pol_keep);
    }
}

template <int KIND>
__device__ UM_EPI_INLINE void epi_chunk_operand(const EpiCtx& E, float acc_scale, uint8_t* a_dst, int* status) {
    const uint64_t pol_keep = policy_evict_last();
    const uint64_t scale2 = pk2(acc_scale, acc_scale);
    const uint32_t t_addr1 = E.t_addr0 + (16u << 16);
    // accumulator column 8i+2tr+cc of a 32-column piece holds logical column 8tr+2i+cc (packing order),
    // so this lane owns 8 consecutive logical columns = one 16-byte row chunk of the next A tile
    const int lcol0 = 8 * E.tr;
    constexpr int kStep = (UM_EPI_WARPS / 4) * UM_PIECE;
    int pc = E.cw * UM_PIECE;
    uint32_t va[16], vb[16];
    if (pc < E.nc) tmem_ld_16x256b_x4(E.t_addr0 + pc, va);            // TMEM reads run one half piece ahead of the math
#pragma unroll 1
    for (int leg = 0; leg < (UM_EARLY_RELEASE ? 2 : 1); ++leg) {
    const int leg_end = (UM_EARLY_RELEASE && leg == 0 && E.rel_bar) ? min(E.nc, UM_SLOT_COLS) : E.nc;
    for (; pc < leg_end; pc += kStep) {
        // the constants of a 32-column block are stored lane-interleaved (float4 j of lane group tr at 4j + tr, j = 2i for
        // pa[i], 2i + 1 for pb[i]): the four lane groups read 64 contiguous bytes - no shared-memory bank conflicts
        float4 pa[4], pb[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            pa[i] = (E.dbg & 8) ? make_float4(0.1f, 0.1f, -1.f, -1.f) : E.par[pc + 8 * i + E.tr];
            if (KIND == EPI_ACT) pb[i] = (E.dbg & 8) ? make_float4(0.5f, 0.5f, 0.5f, 0.5f) : E.par[pc + 8 * i + 4 + E.tr];
        }
        const int kk = E.chunk_col0 + lcol0 + pc;                      // next layer's k index of this lane's 8 values
        // (kk % 32) >> 3 == tr; rows 32q + 16h + 8rr + tq share (row >> 1) & 3 == (tq >> 1) & 3
        uint8_t* dst = a_dst + static_cast<size_t>(kk / UM_KS) * UM_A_STAGE_BYTES + a_tile_off(E.q * 32 + E.tq, E.tr);
        tmem_ld_wait16(va);
        tmem_ld_16x256b_x4(t_addr1 + pc, vb);
        epi_operand_half<KIND>(va, pa, pb, scale2, dst, pol_keep, E.dbg, status);
        tmem_ld_wait16(vb);
        if (pc + kStep < E.nc) tmem_ld_16x256b_x4(E.t_addr0 + pc + kStep, va);
        epi_operand_half<KIND>(vb, pa, pb, scale2, dst + 1024, pol_keep, E.dbg, status);          // rows +16
    }
    if (UM_EARLY_RELEASE && leg == 0 && E.rel_bar) epi_release_first_slot(E);
    }
}

// 32-byte global accesses (sm_100): one lane covers 8 consecutive floats, 4 lanes a full 128-byte line
__device__ __forceinline__ void stg_cs_v8(float* p, const float (&y)[8]) {
    asm volatile("st.global.cs.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
                 :: "l"(p), "f"(y[0]), "f"(y[1]), "f"(y[2]), "f"(y[3]), "f"(y[4]), "f"(y[5]), "f"(y[6]), "f"(y[7]) : "memory");
}
__device__ __forceinline__ void ldg_v8(const float* p, float (&y)[8]) {
    asm volatile("ld.global.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=f"(y[0]), "=f"(y[1]), "=f"(y[2]), "=f"(y[3]), "=f"(y[4]), "=f"(y[5]), "=f"(y[6]), "=f"(y[7]) : "l"(p) : "memory");
}

// store mode of one lane's 8 consecutive output columns
enum { OUT_NONE = -1, OUT_SCALAR = 0, OUT_VEC16 = 1, OUT_VEC32 = 2 };

template <bool TEN_TO, bool ACC>
__device__ __forceinline__ void epi_output_half(const uint32_t (&v)[16], const float4 (&q)[4], float* drow, long long pitch,
                                                int mode, int row_in_q, long long rows_left, int gc, int n_modes, int dbg) {
    if (mode == OUT_NONE) return;
#pragma unroll
    for (int rr = 0; rr < 2; ++rr) {
        float y[8];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            y[2 * i] = fmaf(__uint_as_float(v[4 * i + 2 * rr]), q[i].x, q[i].z);
            y[2 * i + 1] = fmaf(__uint_as_float(v[4 * i + 2 * rr + 1]), q[i].y, q[i].w);
        }
        float* dst = drow + static_cast<long long>(8 * rr) * pitch;
        const bool row_ok = row_in_q + 8 * rr < rows_left;
        if (ACC) {
            // fused sum of log-spectra (e.g. log P_lin + log boost): add what the previous emulator left in `out`
            const float ks = TEN_TO ? 3.3219280948873623f : 1.f;
            if (mode == OUT_VEC32) {
                float p[8];
                ldg_v8(dst, p);
#pragma unroll
                for (int j = 0; j < 8; ++j) y[j] = fmaf(p[j], ks, y[j]);
            } else if (mode == OUT_VEC16) {
                const float4 p0 = *reinterpret_cast<const float4*>(dst), p1 = *reinterpret_cast<const float4*>(dst + 4);
                y[0] = fmaf(p0.x, ks, y[0]); y[1] = fmaf(p0.y, ks, y[1]); y[2] = fmaf(p0.z, ks, y[2]); y[3] = fmaf(p0.w, ks, y[3]);
                y[4] = fmaf(p1.x, ks, y[4]); y[5] = fmaf(p1.y, ks, y[5]); y[6] = fmaf(p1.z, ks, y[6]); y[7] = fmaf(p1.w, ks, y[7]);
            } else if (row_ok) {
#pragma unroll
                for (int j = 0; j < 8; ++j)
                    if (gc + j < n_modes) y[j] = fmaf(dst[j], ks, y[j]);
            }
        }
        if (TEN_TO) {
#pragma unroll
            for (int j = 0; j < 8; ++j) y[j] = ex2_approx(y[j]);
        }
        if (dbg & (4 | 512)) {
            if (y[0] == 1.2345e-30f && y[5] == 3.4e29f) stg_cs_v8(dst, y);
        } else if (mode == OUT_VEC32) {
            stg_cs_v8(dst, y);                 // 8 rows x 128 B per warp instruction: full lines
        } else if (mode == OUT_VEC16) {
            __stcs(reinterpret_cast<float4*>(dst), make_float4(y[0], y[1], y[2], y[3]));
            __stcs(reinterpret_cast<float4*>(dst + 4), make_float4(y[4], y[5], y[6], y[7]));
        } else if (row_ok) {
#pragma unroll
            for (int j = 0; j < 8; ++j)
                if (gc + j < n_modes) __stcs(dst + j, y[j]);
        }
    }
}

// the common case of a half piece - 32-byte stores, every column a mode, every row inside the batch - without a
// branch: both row groups' 16 values go through each step together (FMA, optional 2**y, two 32-byte stores)
template <bool TEN_TO, bool ACC>
// (ACC: p0 / p1 hold the log10 values already in the output rows - loaded by the caller at the top of the piece, so that
// the L2 round trip, which queues behind this warp's own streaming stores, is paid once per piece and under the TMEM wait)
__device__ __forceinline__ void epi_output_half_fast(const uint32_t (&v)[16], const float4 (&q)[4], float* d0, float* d1, int dbg,
                                                     const float (&p0)[8], const float (&p1)[8]) {
    float y0[8], y1[8];
#pragma unroll
    for (int i = 0; i < 4; ++i) {       // packed pairs: (scale0, scale1) x (acc0, acc1) + (shift0, shift1)
        const uint64_t sc = pk2(q[i].x, q[i].y), sh = pk2(q[i].z, q[i].w);
        upk2(ffma2(pk2(__uint_as_float(v[4 * i]), __uint_as_float(v[4 * i + 1])), sc, sh), y0[2 * i], y0[2 * i + 1]);
        upk2(ffma2(pk2(__uint_as_float(v[4 * i + 2]), __uint_as_float(v[4 * i + 3])), sc, sh), y1[2 * i], y1[2 * i + 1]);
    }
    if (ACC) {
        const float ks = TEN_TO ? 3.3219280948873623f : 1.f;
#pragma unroll
        for (int j = 0; j < 8; ++j) { y0[j] = fmaf(p0[j], ks, y0[j]); y1[j] = fmaf(p1[j], ks, y1[j]); }
    }
    if (TEN_TO) {
#pragma unroll
        for (int j = 0; j < 8; ++j) y0[j] = ex2_approx(y0[j]);
#pragma unroll
        for (int j = 0; j < 8; ++j) y1[j] = ex2_approx(y1[j]);
    }
    if (dbg & (4 | 512)) {
        if (y0[0] == 1.2345e-30f && y1[5] == 3.4e29f) stg_cs_v8(d0, y0);
    } else {
        stg_cs_v8(d0, y0);                     // 8 rows x 128 B per warp instruction: full lines
        stg_cs_v8(d1, y1);
    }
}

// output layers: de-standardise (+ 10**y) and store rows of the spectrum.  Columns are packed like the operand
// layers' (a lane owns 8 consecutive logical columns), constants are {scale0, scale1, shift0, shift1}, two columns per
// float4, lane-interleaved (float4 4m + tr of a 32-column block = columns 8tr + 2m, 8tr + 2m + 1).
template <bool TEN_TO, bool ACC>
__device__ __forceinline__ void epi_chunk_output(const EpiCtx& E, float* orow0, long long pitch, int vec_mode,
                                                 long long rows_left, int n_modes) {
    const uint32_t t_addr1 = E.t_addr0 + (16u << 16);
    constexpr int kStep = (UM_EPI_WARPS / 4) * UM_PIECE;
    int pc = E.cw * UM_PIECE;
    // the last chunk of a layer is usually not full: pieces that start beyond the last mode are skipped altogether
    const int nc_valid = min(E.nc, n_modes - E.chunk_col0);
    uint32_t va[16], vb[16];
    if (pc < nc_valid) tmem_ld_16x256b_x4(E.t_addr0 + pc, va);        // TMEM reads run one half piece ahead of the math
#pragma unroll 1
    for (int leg = 0; leg < (UM_EARLY_RELEASE ? 2 : 1); ++leg) {
    const int leg_end = (UM_EARLY_RELEASE && leg == 0 && E.rel_bar) ? min(nc_valid, UM_SLOT_COLS) : nc_valid;
    for (; pc < leg_end; pc += kStep) {
        float4 pp[4];
#pragma unroll
        for (int m = 0; m < 4; ++m) pp[m] = (E.dbg & 8) ? make_float4(1.f, 1.f, 0.f, 0.f) : E.par[(pc >> 1) + 4 * m + E.tr];
        const int gc = E.chunk_col0 + pc + 8 * E.tr;
        float* drow = orow0 + gc;                                       // row 32q + tq of the tile
        // warp-uniform: 32-byte stores, whole tile inside the batch (vec_mode), whole piece inside the spectrum
        const bool fast = vec_mode == OUT_VEC32 && E.chunk_col0 + pc + UM_PIECE <= n_modes;
        // ACC: the log10 values already in the rows.  UM_ACC_HOIST: 0 = loaded where they are used (each load queues behind
        // this warp's streaming stores), 1 = all four rows at the top of the piece, 2 = two at the top, two after the first half
        float prev[UM_ACC_HOIST == 1 ? 4 : 2][8];
        if (ACC && fast && UM_ACC_HOIST >= 1) { ldg_v8(drow, prev[0]); ldg_v8(drow + 8 * pitch, prev[1]); }
        if (ACC && fast && UM_ACC_HOIST == 1) { ldg_v8(drow + 16 * pitch, prev[2]); ldg_v8(drow + 24 * pitch, prev[UM_ACC_HOIST == 1 ? 3 : 1]); }
        tmem_ld_wait16(va);
        tmem_ld_16x256b_x4(t_addr1 + pc, vb);
        if (fast) {
            if (ACC && UM_ACC_HOIST == 0) { ldg_v8(drow, prev[0]); ldg_v8(drow + 8 * pitch, prev[1]); }
            epi_output_half_fast<TEN_TO, ACC>(va, pp, drow, drow + 8 * pitch, E.dbg, prev[0], prev[1]);
            if (ACC && UM_ACC_HOIST != 1) { ldg_v8(drow + 16 * pitch, prev[0]); ldg_v8(drow + 24 * pitch, prev[1]); }
            tmem_ld_wait16(vb);
            if (pc + kStep < nc_valid) tmem_ld_16x256b_x4(E.t_addr0 + pc + kStep, va);
            epi_output_half_fast<TEN_TO, ACC>(vb, pp, drow + 16 * pitch, drow + 24 * pitch, E.dbg, prev[UM_ACC_HOIST == 1 ? 2 : 0], prev[UM_ACC_HOIST == 1 ? 3 : 1]);
            continue;
        }
        const int mode = gc + 8 <= n_modes ? vec_mode : gc < n_modes ? OUT_SCALAR : OUT_NONE;   // OUT_NONE: nothing of this lane's is a mode
        epi_output_half<TEN_TO, ACC>(va, pp, drow, pitch, mode, E.q * 32 + E.tq, rows_left, gc, n_modes, E.dbg);
        tmem_ld_wait16(vb);
        if (pc + kStep < nc_valid) tmem_ld_16x256b_x4(E.t_addr0 + pc + kStep, va);
        epi_output_half<TEN_TO, ACC>(vb, pp, drow + 16 * pitch, pitch, mode, E.q * 32 + E.tq + 16, rows_left, gc, n_modes, E.dbg);
    }
    if (UM_EARLY_RELEASE && leg == 0 && E.rel_bar) epi_release_first_slot(E);
    }
}

// Band-power binning fused into the output epilogue (Planck-lite likelihood, tf_planck2018_lite.py:281-291: window * C_l,
// segment sum): a lane's 8 consecutive multipoles are two groups of four, a group touches at most two bins (the narrowest
// bin is five multipoles wide), so the lane emits four partial sums per row - 16 bytes instead of 32 - in a fixed order;
// cpb200_plik's binning pass adds a bin's partials, again in a fixed order.  Rows leave the SM at half the bytes and the
// tail reads half as much.
__device__ __forceinline__ void epi_binned_half(const uint32_t (&v)[16], const float4 (&pp)[8], float* d0, float* d1, bool ok0, bool ok1) {
    float pa[2][2] = {{0.f, 0.f}, {0.f, 0.f}}, pb[2][2] = {{0.f, 0.f}, {0.f, 0.f}};          // [row group][column group]
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const float c0 = ex2_approx(fmaf(__uint_as_float(v[4 * (j >> 1) + (j & 1)]), pp[j].x, pp[j].y));
        const float c1 = ex2_approx(fmaf(__uint_as_float(v[4 * (j >> 1) + 2 + (j & 1)]), pp[j].x, pp[j].y));
        pa[0][j >> 2] = fmaf(pp[j].z, c0, pa[0][j >> 2]); pb[0][j >> 2] = fmaf(pp[j].w, c0, pb[0][j >> 2]);
        pa[1][j >> 2] = fmaf(pp[j].z, c1, pa[1][j >> 2]); pb[1][j >> 2] = fmaf(pp[j].w, c1, pb[1][j >> 2]);
    }
    if (ok0) __stcs(reinterpret_cast<float4*>(d0), make_float4(pa[0][0], pb[0][0], pa[0][1], pb[0][1]));
    if (ok1) __stcs(reinterpret_cast<float4*>(d1), make_float4(pa[1][0], pb[1][0], pa[1][1], pb[1][1]));
}
// (orow0: partial row of tile row 32q + tq; the buffer must be 16-byte aligned with a pitch that is a multiple of 4 floats)
__device__ __forceinline__ void epi_chunk_binned(const EpiCtx& E, float* orow0, long long pitch, long long rows_left, int n_modes) {
    const uint32_t t_addr1 = E.t_addr0 + (16u << 16);
    const int kStep = (UM_EPI_WARPS / 4) * UM_PIECE;
    int pc = E.cw * UM_PIECE;
    const int nc_valid = min(E.nc, n_modes - E.chunk_col0);
    const int row_in_q = E.q * 32 + E.tq;
    uint32_t va[16], vb[16];
    if (pc < nc_valid) tmem_ld_16x256b_x4(E.t_addr0 + pc, va);
    for (; pc < nc_valid; pc += kStep) {
        float4 pp[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) pp[j] = E.par[pc + 4 * j + E.tr];
        const int gc = E.chunk_col0 + pc + 8 * E.tr;                       // this lane's first multipole: partials at gc / 2
        float* drow = orow0 + (gc >> 1);
        tmem_ld_wait16(va);
        tmem_ld_16x256b_x4(t_addr1 + pc, vb);
        epi_binned_half(va, pp, drow, drow + 8 * pitch, row_in_q < rows_left && gc < n_modes, row_in_q + 8 < rows_left && gc < n_modes);
        tmem_ld_wait16(vb);
        if (pc + kStep < nc_valid) tmem_ld_16x256b_x4(E.t_addr0 + pc + kStep, va);
        epi_binned_half(vb, pp, drow + 16 * pitch, drow + 24 * pitch, row_in_q + 16 < rows_left && gc < n_modes, row_in_q + 24 < rows_left && gc < n_modes);
    }
}

// quadratic form: partial[row] = sum over this warp's columns of acc[row][col] * d[row][col], where d is the fp32 copy
// of the operand rows (the contraction's own A operand).  Columns follow the operand packing: a lane owns 8 consecutive
// logical columns, so d is read 32 bytes at a time, four lanes covering a 128-byte line of a row.
__device__ __forceinline__ void epi_chunk_quadform(const EpiCtx& E, const float* drow0, long long pitch, long long rows_left,
                                                   int n_cols, float* part0, int part_pitch, int lane) {
    const uint32_t t_addr1 = E.t_addr0 + (16u << 16);
    constexpr int kStep = (UM_EPI_WARPS / 4) * UM_PIECE;
    float sum[4] = {0.f, 0.f, 0.f, 0.f};                 // rows tq, tq + 8, tq + 16, tq + 24 of this lane quarter
    const int row_in_q = E.q * 32 + E.tq;
    uint32_t va[16], vb[16];
    for (int pc = E.cw * UM_PIECE; pc < E.nc; pc += kStep) {
        const int gc = E.chunk_col0 + pc + 8 * E.tr;
        tmem_ld_16x256b_x4(E.t_addr0 + pc, va);
        tmem_ld_16x256b_x4(t_addr1 + pc, vb);
        tmem_ld_wait16(va);
        tmem_ld_wait16(vb);
        if (gc + 8 <= n_cols) {
#pragma unroll
            for (int rw = 0; rw < 4; ++rw) {
                if (row_in_q + 8 * rw < rows_left) {
                    float d[8];
                    ldg_v8(drow0 + static_cast<long long>(8 * rw) * pitch + gc, d);
                    const uint32_t (&v)[16] = rw < 2 ? va : vb;
                    const int rr = rw & 1;
#pragma unroll
                    for (int j = 0; j < 8; ++j) sum[rw] = fmaf(__uint_as_float(v[4 * (j >> 1) + 2 * rr + (j & 1)]), d[j], sum[rw]);
                }
            }
        }
    }
#pragma unroll
    for (int rw = 0; rw < 4; ++rw) {
        sum[rw] += __shfl_xor_sync(0xffffffffu, sum[rw], 1);
        sum[rw] += __shfl_xor_sync(0xffffffffu, sum[rw], 2);
        if ((lane & 3) == 0 && row_in_q + 8 * rw < rows_left) part0[static_cast<long long>(8 * rw) * part_pitch] = sum[rw];
    }
}

// ---------------------------------------------------------------------------------------------
// the kernel
// ---------------------------------------------------------------------------------------------
// PROF = true instantiates the per-role cycle counters (cpb200_engine_debug_counters); the production
// instantiation carries no clock reads on the single-thread issue paths
// SPLIT = true instantiates the small-batch variant (UmmaParams::out_split > 1); the throughput kernel carries none of it
// MODE = 1 (UM_MODE_ACC) instantiates the output epilogues that add the log10 values already in the output rows (fused sums
// of emulators), MODE = 2 (UM_MODE_BINNED) the ones that write partial band powers; the plain kernel (MODE = 0) carries
// neither (its register allocation is the epilogue's)
enum { UM_MODE_PLAIN = 0, UM_MODE_ACC = 1, UM_MODE_BINNED = 2 };
template <bool PROF, bool SPLIT = false, int MODE = UM_MODE_PLAIN>
__global__ void __cluster_dims__(UM_CLUSTER, 1, 1) __launch_bounds__(UM_THREADS, 1)
fused_mlp_umma_kernel(const __grid_constant__ UmmaParams P)
{
    auto mbar_wait = [](uint32_t bar, uint32_t parity, int* error_flag, int code) -> long long {
        if (PROF) return cpb::mbar_wait(bar, parity, error_flag, code);
        if (!mbar_try_wait(bar, parity)) mbar_wait_slow(bar, parity, error_flag, code);
        return 0;
    };
    auto now = []() -> long long { return PROF ? clock64() : 0; };
    const long long t_kernel0 = clock64();
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;
    uint8_t* sbase = smem_raw + (base - raw);

    // broadcast values are warp-uniform for the compiler: the issuing roles below then keep their addresses
    // and descriptors in uniform registers instead of re-broadcasting them before every copy / MMA
    const int warp = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0);
    const int lane = threadIdx.x & 31;

    const uint32_t bars = base + UM_SMEM_BARS;
    constexpr int B_ACC = 2 * UM_STAGES, B_AREADY = B_ACC + 2 * UM_TMEM_SLOTS, B_A0 = B_AREADY + 4, B_PAR = B_A0 + 4 * UM_MAX_MODELS;
    auto bar_full = [&](int s) { return bars + 8u * s; };
    auto bar_empty = [&](int s) { return bars + 8u * (UM_STAGES + s); };
    auto bar_accfull = [&](int sl) { return bars + 8u * (B_ACC + sl); };                    // per TMEM slot (a job signals its first slot)
    auto bar_accempty = [&](int sl) { return bars + 8u * (B_ACC + UM_TMEM_SLOTS + sl); };   // leader only: both CTAs' epilogues arrive
    auto bar_aready = [&](int ts) { return bars + 8u * (B_AREADY + ts); };                  // ts = 2 * (pair number & 1) + tile
    auto bar_a0ready = [&](int mt) { return bars + 8u * (B_A0 + mt); };                     // mt = 2 * model + tile
    auto bar_a0free = [&](int mt) { return bars + 8u * (B_A0 + 2 * UM_MAX_MODELS + mt); };
    auto bar_parfull = [&](int b) { return bars + 8u * (B_PAR + b); };
    auto bar_parempty = [&](int b) { return bars + 8u * (B_PAR + UM_PAR_SLOTS + b); };
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(sbase + UM_SMEM_TMEMPTR);

    if (threadIdx.x == 0) {
        // the leader's "full" barrier of a stage completes when its own copies have landed (arrive.expect_tx by
        // its loader) AND the peer's relay has reported the peer's copies (one remote arrive): one wait per stage
        const uint32_t full_count = (UM_CLUSTER > 1 && cluster_ctarank() == 0) ? 2 : 1;
        for (int s = 0; s < UM_STAGES; ++s) { mbar_init(bar_full(s), full_count); mbar_init(bar_empty(s), 1); }
        for (int b = 0; b < UM_TMEM_SLOTS; ++b) { mbar_init(bar_accfull(b), 1); mbar_init(bar_accempty(b), UM_CLUSTER * UM_EPI_WARPS); }
        for (int b = 0; b < UM_PAR_SLOTS; ++b) { mbar_init(bar_parfull(b), 1); mbar_init(bar_parempty(b), UM_EPI_WARPS); }
        for (int ts = 0; ts < 4; ++ts) mbar_init(bar_aready(ts), UM_EPI_WARPS);
        for (int mt = 0; mt < 2 * UM_MAX_MODELS; ++mt) { mbar_init(bar_a0ready(mt), 1); mbar_init(bar_a0free(mt), 1); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        // pair-wide allocation: the same warp of both CTAs issues it; each CTA gets the same 512 columns
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;"
                     ::"r"(smem_u32(tmem_ptr_smem)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    if (UM_CLUSTER > 1) cluster_sync_all();      // peers' barriers are initialised before any remote arrive / multicast
    tc_fence_after();
    const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_ptr_smem, 0);
    const uint32_t cta_rank = UM_CLUSTER > 1 ? __shfl_sync(0xffffffffu, cluster_ctarank(), 0) : 0;
    const int out_split = SPLIT ? P.out_split : 1, split_id = SPLIT ? static_cast<int>(blockIdx.x / UM_CLUSTER) % out_split : 0;
    const int out_gemm = SPLIT ? P.n_gemms - 1 : -1;       // the contraction whose chunks are dealt out (single-emulator launches only)
    const uint32_t a0_area = 2u * P.n_models * P.a0_bytes;  // layer-0 operand tiles: one per (emulator, tile)
    const long long pg_first = (blockIdx.x / UM_CLUSTER) / out_split, pg_step = (gridDim.x / UM_CLUSTER) / out_split;
    // pair-groups this cluster walks; iteration `it` runs the table's "cur" entries for pair-group it-1 and its "next"
    // entries for pair-group it (layer 0 only, or the whole trunk when the output phase is overlapped): work of the next
    // pair fills the phases of the current one in which the tensor pipe would wait for epilogues
    const long long n_my_pg = P.n_pairgroups > pg_first ? (P.n_pairgroups - pg_first + pg_step - 1) / pg_step : 0;
    const int n_jobs = P.n_jobs;

    uint8_t* cta_scratch = P.scratch + static_cast<unsigned long long>(blockIdx.x) * P.scratch_cta_bytes;
    // CTA-private scratch: [2 layer-0 operand tiles per emulator][UM_ACT_BUFS rotating activation buffers][4 parity buffers].
    // Rotating buffers: with the job order X.g, Y.g, X.g+1, ... at most three 128 x K tiles are live at any time
    // (X.in, X.out, Y.in - or Y.in, Y.out, X.out); the host numbers them so that step n = 2(g-1) + t reads buffer 2n and
    // writes buffer 2n + 1 (mod 3), plus a per-pair offset (rot_mul) that lets the next pair's layer 0, which runs inside
    // this pair's output phase, land in the buffer that is free then.  3 x 256 KB per CTA keeps the scratch at 111 MB.
    // Parity buffers hold the operand of an overlapped output phase: written by pair p's trunk, read during pair p+1's.
    // (pm3 = pair number mod 3, ppar = its parity: kept as small ints so that no 64-bit division sits on any role's path)
    auto buf_addr = [&](uint32_t code, int pm3, int ppar) -> uint8_t* {
        uint32_t off;
        if (code & UM_BUF_ROT) off = a0_area + (((code & 0xf) + pm3 * P.rot_mul) % UM_ACT_BUFS) * P.abuf_bytes;
        else if (code & UM_BUF_PAR) off = a0_area + UM_ACT_BUFS * P.abuf_bytes + (2 * ppar + (code & 1)) * P.pbuf_bytes;
        else off = (code & 0xf) * P.a0_bytes;                     // UM_BUF_A0 + 2 * model + tile
        return cta_scratch + off;
    };

// every role walks the same entry sequence
#define UM_FOR_EACH_ENTRY(...)                                                                         \
    for (long long it = 0; it <= n_my_pg; ++it) {                                                      \
        const bool has_cur = it >= 1, has_next = it < n_my_pg;                                         \
        const int it_m3 = static_cast<int>(it % 3), it_par = static_cast<int>(it & 1);                 \
        for (int ji = 0; ji < n_jobs; ++ji) {                                                          \
            const uint32_t jw = P.jobs[ji];                                                            \
            const int g = jw & 15, t = (jw >> 4) & 1, c = (jw >> 5) & 127;                             \
            const bool is_next = (jw >> 12) & 1;                                                       \
            if (is_next ? !has_next : !has_cur) continue;                                              \
            if (SPLIT && g == out_gemm && c % out_split != split_id) continue;   /* another cluster's columns */ \
            const int s0 = (jw >> 13) & 63, ns = (jw >> 19) & 63;                                      \
            const bool first = (jw >> 25) & 1, last = (jw >> 26) & 1;                                  \
            const int slot = (jw >> 27) & 3, two = (jw >> 29) & 1;                                     \
            /* rows beyond n_rows are padding (computed, never stored) */                              \
            const long long pidx = is_next ? it : it - 1;      /* this cluster's running pair number */  \
            const long long pair = (pg_first + pidx * pg_step) * UM_CLUSTER + cta_rank;                \
            /* a pair-group whose second tiles hold no row at all (the leader's tile 1 starts beyond the batch, so  \
               the peer's tiles do too): every role skips the t = 1 jobs - a batch of up to 128 rows, a sampler's   \
               single point, then walks one tile through the network instead of two (it is a cluster's last group) */ \
            if (t == 1 && ((pg_first + pidx * pg_step) * UM_CLUSTER * 2 + 1) * UM_TILE_M >= P.n_rows) continue; \
            const int ppar = is_next ? it_par : it_par ^ 1;            /* parity and residue mod 3 of the pair number */ \
            const int pm3 = is_next ? it_m3 : (it_m3 + 2) % 3;                                         \
            const int ts = 2 * ppar + t;                                                               \
            (void)pair; (void)c; (void)pidx; (void)s0; (void)ns; (void)first; (void)last; (void)slot; (void)two; (void)ts; (void)pm3; \
            __VA_ARGS__                                                                                \
        }                                                                                              \
    }

    if (warp == 0) {
        // ===================== loader: bulk copies into the stage ring =====================
        // The whole warp walks the loop (warp-uniform control flow and values); one elected lane issues.
        {
            long long w_empty = 0, w_aready = 0; const long long t_begin = now();
            int stage = 0; uint32_t ph_empty = 1;                 // fresh "empty" barriers pass at parity 1
            uint32_t ph_aready = 0, ph_a0ready = 0;               // one phase bit per barrier
            uint32_t ph_parempty = (1u << UM_PAR_SLOTS) - 1;      // one phase bit per slot
            const uint64_t pol_keep = policy_evict_last(), pol_stream = policy_evict_first();
            int pb = 0;
            UM_FOR_EACH_ENTRY({
                const UmmaGemm& L = P.g[g];
                const uint32_t b_bytes = 2u * L.nc * UM_KS * 2u;
                const long long j_t0 = now(), j_we0 = w_empty, j_wa0 = w_aready;
                const int mt = 2 * L.model + t;
                if (first && c == (g == out_gemm ? split_id : 0) && !(L.layer0 && P.ext_a0)) {   // first chunk this cluster runs of (g, t): the whole K extent of the tile's operand must be written
                    if (L.layer0) { w_aready += mbar_wait(bar_a0ready(mt), (ph_a0ready >> mt) & 1, P.error_flag, 101); ph_a0ready ^= 1u << mt; }
                    else { w_aready += mbar_wait(bar_aready(ts), (ph_aready >> ts) & 1, P.error_flag, 102); ph_aready ^= 1u << ts; }
                }
                const uint8_t* a_src = (L.layer0 && P.ext_a0) ? P.ext_a0 + static_cast<size_t>(pair * 2 + t) * P.a0_bytes
                                                            : buf_addr(P.jobs_x[ji] & 0xff, pm3, ppar);
                a_src += static_cast<size_t>(s0) * UM_A_STAGE_BYTES;
                if (last) {   // this chunk's per-column epilogue constants -> shared memory (ring in the order the jobs COMPLETE)
                    w_empty += mbar_wait(bar_parempty(pb), (ph_parempty >> pb) & 1, P.error_flag, 104); ph_parempty ^= 1u << pb;
                    // operand layers: one float4 per column; output layers: {scale, shift} pairs, two columns per
                    // float4, in two variants (plain, then pre-multiplied by log2(10) for ten_to)
                    const bool is_binned = MODE == UM_MODE_BINNED && L.epi_kind == EPI_OUT && P.m[L.model].binned;
                    const bool is_out = L.epi_kind == EPI_OUT && !is_binned;
                    const uint32_t par_bytes = is_out ? L.nc * 8u : L.nc * 16u;
                    const float4* par_src = is_binned ? L.epi_b + c * L.nc
                                          : is_out ? L.epi + (P.m[L.model].ten_to ? L.n_chunks * (L.nc / 2) : 0) + c * (L.nc / 2) : L.epi + c * L.nc;
                    if (elect_one()) {
                        mbar_arrive_expect_tx(bar_parfull(pb), par_bytes);
                        bulk_g2s(base + UM_SMEM_EPIPAR + pb * (UM_MAX_NC * 16), par_src, par_bytes, bar_parfull(pb));
                    }
                    if (++pb == UM_PAR_SLOTS) pb = 0;
                }
                // this CTA's half of each weight stage ([hi | lo] of its nc/2 columns)
                const uint8_t* b_src = reinterpret_cast<const uint8_t*>(L.wpack) + (static_cast<size_t>(c) * L.k_stages + s0) * b_bytes
                                     + cta_rank * (b_bytes / 2);
                // an activation tile is dead after its last chunk has streamed it
                const uint64_t pol_a = c == L.n_chunks - 1 ? pol_stream : pol_keep;
                const uint32_t tx_bytes = UM_A_STAGE_BYTES + b_bytes / 2;
                for (int s = 0; s < ns; ++s) {
                    w_empty += mbar_wait(bar_empty(stage), ph_empty, P.error_flag, 103);
                    if (PROF && P.counters && blockIdx.x == 0 && it == UM_JOB_TRACE_ITERATION && s < UM_TRACE_STAGES && lane == 0)
                        P.counters[(gridDim.x + UM_MAX_JOBS) * UM_CNT_SLOTS + (ji * UM_TRACE_STAGES + s) * 2] = now();
                    if (PROF && (P.debug_flags & 128)) {      // ablation: no operand traffic (the stage is only handed on)
                        if (elect_one()) mbar_arrive(bar_full(stage));
                    } else
                    if (elect_one()) {
                        const uint32_t dst = base + UM_SMEM_STAGES + stage * UM_STAGE_BYTES;
                        mbar_arrive_expect_tx(bar_full(stage), tx_bytes);
                        bulk_g2s_hint(dst, a_src, UM_A_STAGE_BYTES, bar_full(stage), pol_a);
                        bulk_g2s_hint(dst + UM_A_STAGE_BYTES, b_src, b_bytes / 2, bar_full(stage), pol_keep);
                    }
                    a_src += UM_A_STAGE_BYTES;
                    b_src += b_bytes;
                    if (++stage == UM_STAGES) { stage = 0; ph_empty ^= 1; }
                }
                if (PROF && P.counters && blockIdx.x == 0 && lane == 0) {
                    long long* jc = P.counters + (gridDim.x + ji) * UM_CNT_SLOTS;
                    jc[UM_JOB_LOAD_WAIT_AREADY] += w_aready - j_wa0; jc[UM_JOB_LOAD_WAIT_EMPTY] += w_empty - j_we0;
                    if (it == UM_JOB_TRACE_ITERATION) { jc[UM_JOB_T_LOAD0] = j_t0; jc[UM_JOB_T_LOAD1] = now(); }
                }
            })
            if (PROF && P.counters && lane == 0) {
                long long* cnt = P.counters + blockIdx.x * UM_CNT_SLOTS;
                cnt[UM_CNT_LOAD_WAIT_EMPTY] = w_empty; cnt[UM_CNT_LOAD_WAIT_AREADY] = w_aready;
                cnt[UM_CNT_LOAD_TOTAL] = now() - t_begin;
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer (leader CTA) / stage relay (peer CTA) =====================
        // whole-warp, warp-uniform loops as in the loader; one elected lane issues
        if (cta_rank != 0) {
            // the peer's operands live in its own shared memory: tell the leader when each stage has landed
            int stage = 0; uint32_t ph_full = 0;
            const uint32_t leader_full0 = map_to_cta(bar_full(0), 0);     // the leader's full barriers
            UM_FOR_EACH_ENTRY({
                for (int i = 0; i < ns; ++i) {
                    mbar_wait(bar_full(stage), ph_full, P.error_flag, 203);
                    if (elect_one()) mbar_arrive_remote_relaxed(leader_full0 + 8u * stage);
                    if (++stage == UM_STAGES) { stage = 0; ph_full ^= 1; }
                }
            })
        } else {
            long long w_full = 0, w_accempty = 0; const long long t_begin = now();
            int stage = 0; uint32_t ph_full = 0;
            uint32_t ph_accempty = (1u << UM_TMEM_SLOTS) - 1;     // one phase bit per TMEM slot
            // shared-memory descriptors of ring stage 0; a stage advances the 14-bit address field by its size >> 4
            const uint64_t a_hi0 = umma_desc_a(base + UM_SMEM_STAGES);
            const uint64_t a_lo0 = umma_desc_a(base + UM_SMEM_STAGES + UM_A_HALF_BYTES);
            const uint64_t b_hi0 = umma_desc(base + UM_SMEM_STAGES + UM_A_STAGE_BYTES);
            long long aligned_it = 0;
            const long long align_last = P.align_every > 0 ? (P.n_pairgroups / (gridDim.x / UM_CLUSTER)) : 0;   // iterations every cluster runs
            UM_FOR_EACH_ENTRY({
                if (PROF && P.align_every > 0 && P.align_counter && it != aligned_it && it % P.align_every == 0 && it <= align_last && out_split == 1) {
                    aligned_it = it;
                    if (lane == 0) {
                        const unsigned target = static_cast<unsigned>((it / P.align_every) * (gridDim.x / UM_CLUSTER));
                        atomicAdd(P.align_counter, 1u);
                        const long long t0 = clock64();
                        while (*reinterpret_cast<volatile unsigned*>(P.align_counter) < target) {
                            if (clock64() - t0 > CPB_WATCHDOG_CYCLES) { if (P.error_flag) atomicExch(P.error_flag, 210); __threadfence_system(); __trap(); }
                        }
                    }
                    __syncwarp();
                }
                const UmmaGemm& L = P.g[g];
                const uint32_t idesc = umma_idesc(L.nc);
                const uint64_t b_lo_off = static_cast<uint64_t>((static_cast<uint32_t>(L.nc / 2) * UM_KS * 2u) >> 4);   // (hi -> lo half tile) >> 4
                const long long j_t0 = now(), j_wf0 = w_full, j_wa0 = w_accempty;
                if (first) {
                    // (a cluster-scope acquire would cost an L1 invalidate per wait on the issue path; the peer
                    // only hands back TMEM it has finished reading - tcgen05.wait::ld precedes its arrive)
                    w_accempty += mbar_wait(bar_accempty(slot), (ph_accempty >> slot) & 1, P.error_flag, 201); ph_accempty ^= 1u << slot;
                    if (two) { w_accempty += mbar_wait(bar_accempty(slot + 1), (ph_accempty >> (slot + 1)) & 1, P.error_flag, 201); ph_accempty ^= 1u << (slot + 1); }
                    tc_fence_after();
                }
                const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(slot) * UM_SLOT_COLS;
                for (int s = 0; s < ns; ++s) {
                    w_full += mbar_wait(bar_full(stage), ph_full, P.error_flag, 202);
                    if (PROF && P.counters && blockIdx.x == 0 && it == UM_JOB_TRACE_ITERATION && s < UM_TRACE_STAGES && lane == 0)
                        P.counters[(gridDim.x + UM_MAX_JOBS) * UM_CNT_SLOTS + (ji * UM_TRACE_STAGES + s) * 2 + 1] = now();
                    tc_fence_after();
                    if (PROF && (P.debug_flags & 64)) {       // ablation: no tensor-core work at all (stages and accumulators are only handed on)
                        if (elect_one()) umma_commit_2cta(bar_empty(stage));
                    } else
                    if (elect_one()) {
                        const uint64_t adv = static_cast<uint64_t>(stage) * (UM_STAGE_BYTES >> 4);
                        const uint64_t a_hi = a_hi0 + adv, a_lo = a_lo0 + adv, b_hi = b_hi0 + adv, b_lo = b_hi + b_lo_off;
#pragma unroll
                        for (int kk = 0; kk < UM_KS / 16; ++kk) {
                            const uint64_t adv_b = static_cast<uint64_t>((kk * 256) >> 4);  // B: 2 core matrices along K
                            const uint64_t adv_a = static_cast<uint64_t>((kk * 32) >> 4);   // A: 16 k = 32 B inside the swizzled row
                            umma_f16_2cta(d_tmem, a_hi + adv_a, b_hi + adv_b, idesc, ((s0 + s) | kk) != 0);
                            umma_f16_2cta(d_tmem, a_lo + adv_a, b_hi + adv_b, idesc | UM_IDESC_NEG_A, 1u);   // the tile holds -lo
                            umma_f16_2cta(d_tmem, a_hi + adv_a, b_lo + adv_b, idesc, 1u);
                        }
                        umma_commit_2cta(bar_empty(stage));       // frees the stage in both CTAs
                    }
                    if (++stage == UM_STAGES) { stage = 0; ph_full ^= 1; }
                }
                if (last && elect_one()) {
                    umma_commit_2cta(bar_accfull(slot));          // both CTAs' epilogues may drain their 128 rows
                    if (L.layer0 && c == L.n_chunks - 1) umma_commit_2cta(bar_a0free(2 * L.model + t));   // layer-0 operand of (emulator, tile) consumed
                }
                if (PROF && P.counters && blockIdx.x == 0 && lane == 0) {
                    long long* jc = P.counters + (gridDim.x + ji) * UM_CNT_SLOTS;
                    const long long j_t1 = now();
                    jc[UM_JOB_MMA_WAIT_FULL] += w_full - j_wf0; jc[UM_JOB_MMA_WAIT_ACCEMPTY] += w_accempty - j_wa0;
                    jc[UM_JOB_MMA_BUSY] += j_t1 - j_t0; jc[UM_CNT_SLOTS - 1] = jw;
                    if (it == UM_JOB_TRACE_ITERATION) { jc[UM_JOB_T_MMA0] = j_t0; jc[UM_JOB_T_MMA1] = j_t1; }
                }
            })
            if (PROF && P.counters && lane == 0) {
                long long* cnt = P.counters + blockIdx.x * UM_CNT_SLOTS;
                cnt[UM_CNT_MMA_WAIT_FULL] = w_full; cnt[UM_CNT_MMA_WAIT_ACCEMPTY] = w_accempty;
                cnt[UM_CNT_MMA_TOTAL] = now() - t_begin;
            }
        }
    } else if (warp == 2) {
        // ===================== parameter standardiser: builds the layer-0 A operand =====================
        uint32_t ph_a0free = (1u << (2 * UM_MAX_MODELS)) - 1;     // one phase bit per (emulator, tile)
        long long w_in = 0; const long long t_begin = now();
        // the tiles are consumed in the order (pair, emulator, tile): the job table runs the emulators of a launch one
        // after the other on each pair of tiles
        for (long long pgm = 0; pgm < (P.ext_a0 ? 0 : n_my_pg * P.n_models); ++pgm) {      // (prebuilt operand tiles: nothing to do)
            const long long pgi = pgm / P.n_models;
            const int mi = static_cast<int>(pgm % P.n_models);
            const UmmaModel& M = P.m[mi];
            const int np = M.n_parameters;
            const long long pair = (pg_first + pgi * pg_step) * UM_CLUSTER + cta_rank;      // rows beyond n_rows are padding
            for (int t = 0; t < 2; ++t) {
                const int mt = 2 * mi + t;
                w_in += mbar_wait(bar_a0free(mt), (ph_a0free >> mt) & 1, P.error_flag, 301); ph_a0free ^= 1u << mt;
                const long long row0 = (pair * 2 + t) * UM_TILE_M;
                uint8_t* dst = cta_scratch + static_cast<size_t>(mt) * P.a0_bytes;
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
                                const float raw_v = (grow < P.n_rows) ? M.params[grow * np + k] : M.pmean[k];
                                v = (raw_v - M.pmean[k]) / M.pstd[k] * UM_INPUT_SCALE;
                                // beyond +-64 sigma (or non-finite) the split would saturate: flag the batch
                                if (UM_RANGE_CHECK && !(fabsf(v) < UM_F16_MAX) && P.error_flag) *reinterpret_cast<volatile int*>(P.error_flag + 1) = 1;
                            }
                            x[u] = v;
                        }
                        uint4 hi, lo;
                        split2n(x[0], x[1], hi.x, lo.x); split2n(x[2], x[3], hi.y, lo.y);
                        split2n(x[4], x[5], hi.z, lo.z); split2n(x[6], x[7], hi.w, lo.w);
                        const size_t off = static_cast<size_t>(k0 / UM_KS) * UM_A_STAGE_BYTES + a_tile_off(r, (k0 % UM_KS) >> 3);
                        *reinterpret_cast<uint4*>(dst + off) = hi;
                        *reinterpret_cast<uint4*>(dst + off + UM_A_HALF_BYTES) = lo;
                    }
                }
                __threadfence();
                fence_proxy_async();
                __syncwarp();
                if (lane == 0) mbar_arrive(bar_a0ready(mt));
            }
        }
        if (PROF && P.counters && lane == 0) {
            long long* cnt = P.counters + blockIdx.x * UM_CNT_SLOTS;
            cnt[UM_CNT_IN_WAIT] = w_in; cnt[UM_CNT_IN_TOTAL] = now() - t_begin;
        }
    } else if (warp >= UM_EPI_WARP0) {
        // ===================== epilogue warps =====================
        // They take the jobs in the order in which the table COMPLETES them (the entries flagged `last`).
        const int ew = warp - UM_EPI_WARP0;          // 0..7
        const int q = warp & 3;                      // TMEM lane quarter this warp may read
        const int cw = ew >> 2;                      // which of the lane quarter's warps: takes every (UM_EPI_WARPS/4)-th 32-column piece
        const float4* epipar = reinterpret_cast<const float4*>(sbase + UM_SMEM_EPIPAR);
        const int tq = lane >> 2, tr = lane & 3;     // fragment coordinates: rows tq, tq+8 (+16 for the 2nd load); cols 2*tr, 2*tr+1
        uint32_t ph_accfull = 0, ph_parfull = 0;     // one phase bit per TMEM slot / constant slot
        long long w_accfull = 0, c_act = 0, c_out = 0; const long long t_begin = now();
        int pb = 0;
        EpiCtx E;
        E.q = q; E.tq = tq; E.tr = tr; E.cw = cw; E.lane = lane;
#ifdef UM_PROF_NO_DBG
        E.dbg = 0;                               // profiling build whose epilogue loops are the production ones (no ablation paths)
#else
        E.dbg = PROF ? P.debug_flags : 0;
#endif
        E.leader = cta_rank == 0;
        const uint32_t accempty0 = E.leader ? bar_accempty(0) : map_to_cta(bar_accempty(0), 0);     // the leader's MMA issuer owns the TMEM hand-back
        UM_FOR_EACH_ENTRY({
            if (!last) continue;
            const UmmaGemm& L = P.g[g];
            E.nc = L.nc;
            // quadratic-form chunks hand both slots back at the end
            E.rel_bar = (UM_EARLY_RELEASE && two && L.epi_kind != EPI_QUADFORM && !(PROF && (P.debug_flags & 3))) ? accempty0 + 8u * slot : 0u;
            const long long row0 = (pair * 2 + t) * UM_TILE_M;
            E.par = epipar + pb * UM_MAX_NC;                  // filled by the loader's bulk copy
            const long long j_w0 = w_accfull;
            w_accfull += mbar_wait(bar_parfull(pb), (ph_parfull >> pb) & 1, P.error_flag, 402); ph_parfull ^= 1u << pb;
            w_accfull += mbar_wait(bar_accfull(slot), (ph_accfull >> slot) & 1, P.error_flag, 401); ph_accfull ^= 1u << slot;
            tc_fence_after();
            const long long t_job = now();
            E.t_addr0 = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(slot * UM_SLOT_COLS);
            E.chunk_col0 = c * L.nc;
            uint8_t* const o_buf = buf_addr(P.jobs_x[ji] >> 8, pm3, ppar);
            if (PROF && (P.debug_flags & 3)) {
                if (P.debug_flags & 2) {           // TMEM reads only
                    uint32_t va[16];
                    for (int pc = cw * UM_PIECE; pc < L.nc; pc += (UM_EPI_WARPS / 4) * UM_PIECE) {
                        tmem_ld_16x256b_x4(E.t_addr0 + pc, va); tmem_ld_wait16(va);
                        tmem_ld_16x256b_x4(E.t_addr0 + (16u << 16) + pc, va); tmem_ld_wait16(va);
                    }
                }
            } else
            if (L.epi_kind == EPI_ACT) epi_chunk_operand<EPI_ACT>(E, L.acc_scale, o_buf, P.error_flag);
            else if (L.epi_kind == EPI_AFFINE) epi_chunk_operand<EPI_AFFINE>(E, 1.f, o_buf, P.error_flag);
            else if (L.epi_kind == EPI_QUADFORM)
                epi_chunk_quadform(E, P.qf_d + (row0 + q * 32 + tq) * P.qf_pitch, P.qf_pitch, P.n_rows - row0, P.qf_cols,
                                   P.qf_partial + (row0 + q * 32 + tq) * (2 * L.n_chunks) + 2 * c + cw, 2 * L.n_chunks, lane);
            else {
                const UmmaModel& M = P.m[L.model];
                // 32-byte (16-byte) output stores need a 32-byte (16-byte) aligned base and a row pitch that is a multiple of 8 (4)
                // floats - and a tile that lies wholly inside the batch
                const int vec_mode = row0 + UM_TILE_M > P.n_rows ? OUT_SCALAR
                                   : ((reinterpret_cast<uintptr_t>(M.out) & 31) | (static_cast<uintptr_t>(M.out_pitch) & 7)) == 0 ? OUT_VEC32
                                   : ((reinterpret_cast<uintptr_t>(M.out) & 15) | (static_cast<uintptr_t>(M.out_pitch) & 3)) == 0 ? OUT_VEC16 : OUT_SCALAR;
                // row handled by fragment slot (h, rr): row0 + 32q + 16h + 8rr + tq
                float* orow0 = M.out + (row0 + q * 32 + tq) * M.out_pitch;
                if (MODE == UM_MODE_BINNED && M.binned) {
                    epi_chunk_binned(E, orow0, M.out_pitch, P.n_rows - row0, M.n_modes);
                } else if (MODE == UM_MODE_ACC && M.accumulate) {
                    if (M.ten_to) epi_chunk_output<true, true>(E, orow0, M.out_pitch, vec_mode, P.n_rows - row0, M.n_modes);
                    else epi_chunk_output<false, true>(E, orow0, M.out_pitch, vec_mode, P.n_rows - row0, M.n_modes);
                } else {
                    if (M.ten_to) epi_chunk_output<true, false>(E, orow0, M.out_pitch, vec_mode, P.n_rows - row0, M.n_modes);
                    else epi_chunk_output<false, false>(E, orow0, M.out_pitch, vec_mode, P.n_rows - row0, M.n_modes);
                }
            }
            // accumulator drained (every lane has passed tcgen05.wait::ld): hand the rest of its TMEM back to the MMA
            // issuer first - the remainder of this job's tail does not touch TMEM
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
                if (E.leader) {
                    if (!E.rel_bar) mbar_arrive_relaxed(accempty0 + 8u * slot);
                    if (two) mbar_arrive_relaxed(accempty0 + 8u * (slot + 1));
                } else {                                                          // TMEM reads are complete (wait::ld)
                    if (!E.rel_bar) mbar_arrive_remote_relaxed(accempty0 + 8u * slot);
                    if (two) mbar_arrive_remote_relaxed(accempty0 + 8u * (slot + 1));
                }
                mbar_arrive_relaxed(bar_parempty(pb));
            }
            if (++pb == UM_PAR_SLOTS) pb = 0;
            if (L.epi_kind <= EPI_AFFINE && c == L.n_chunks - 1) {     // operand-producing kinds
                // the next layer's loader may only read the scratch once every chunk of this layer is
                // written: one fence per thread before the last chunk's arrive covers all its stores
                UM_SCRATCH_FENCE(); fence_proxy_async_global();
                __syncwarp();
                if (lane == 0) mbar_arrive(bar_aready(ts));
            }
            if (UM_DISCARD_DEAD_TILES && P.discard_ok && !L.layer0 && c == L.n_chunks - 1 && !(PROF && (P.debug_flags & 256))) {
                // This was the last chunk to stream the tile's operand: every copy of it has landed and been consumed
                // (its MMAs are complete), so its L2 lines are dead - drop them instead of letting them be written
                // back to HBM on eviction.  The buffer is rewritten soon, possibly while slower warps
                // are still in this job, so a warp only drops the lines it will itself rewrite (program order then
                // keeps discard before store): with chunk widths that are multiples of 64 (discard_ok) warp (q, cw)
                // writes exactly the rows of lane quarter q in the K stages of parity cw - 32 lines per stage.
                const uint8_t* dead = buf_addr(P.jobs_x[ji] & 0xff, pm3, ppar) + (lane >> 4) * UM_A_HALF_BYTES + q * 2048 + (lane & 15) * 128;
                for (int sd = cw; sd < L.k_stages; sd += UM_EPI_WARPS / 4) l2_discard_line(dead + static_cast<size_t>(sd) * UM_A_STAGE_BYTES);
            }
            if (L.epi_kind >= EPI_OUT) c_out += now() - t_job; else c_act += now() - t_job;
            if (PROF && P.counters && blockIdx.x == 0 && ew == 0 && lane == 0) {
                long long* jc = P.counters + (gridDim.x + ji) * UM_CNT_SLOTS;
                const long long j_t1 = now();
                jc[UM_JOB_EPI_WAIT] += w_accfull - j_w0; jc[UM_JOB_EPI_BUSY] += j_t1 - t_job;
                if (it == UM_JOB_TRACE_ITERATION) { jc[UM_JOB_T_EPI0] = t_job; jc[UM_JOB_T_EPI1] = j_t1; }
            }
        })
        if (PROF && P.counters && ew == 0 && lane == 0) {
            long long* cnt = P.counters + blockIdx.x * UM_CNT_SLOTS;
            cnt[UM_CNT_EPI_WAIT_ACCFULL] = w_accfull; cnt[UM_CNT_EPI_ACT] = c_act; cnt[UM_CNT_EPI_OUT] = c_out;
            cnt[UM_CNT_EPI_TOTAL] = now() - t_begin;
        }
    }
#undef UM_FOR_EACH_ENTRY

    tc_fence_before();
    __syncthreads();
    if (UM_CLUSTER > 1) cluster_sync_all();      // no CTA leaves while a peer can still multicast into it / arrive on it
    if (threadIdx.x == 0 && P.max_cycles) atomicMax(P.max_cycles, static_cast<unsigned long long>(clock64() - t_kernel0));
    if (warp == 2) {
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    }
}

}  // namespace cpb
