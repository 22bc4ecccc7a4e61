// C ABI of the B200 CosmoPower engine (see include/cpb200.h).  Host-side logic only:
// model validation, weight packing/upload, scratch management, kernel launches and the pipelined
// host<->device path.  No CPU compute fallback exists: without an sm_100 device every entry point
// that would compute returns CPB200_E_NO_DEVICE.
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <pthread.h>
#include <sched.h>
#include <unistd.h>

#include <algorithm>
#include <atomic>
#include <cctype>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <string>
#include <thread>
#include <type_traits>
#include <vector>

#include "../../include/cpb200.h"
#include "simt_kernels.cuh"
#include "umma_kernel.cuh"
#include "plik_kernels.cuh"

namespace {

thread_local std::string g_err;

int fail(int code, const char* fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_err = buf;
    return code;
}

#define CU(call)                                                                                  \
    do {                                                                                          \
        cudaError_t _e = (call);                                                                  \
        if (_e != cudaSuccess)                                                                    \
            return fail(CPB200_E_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(_e),   \
                        __FILE__, __LINE__);                                                      \
    } while (0)

struct DeviceGuard {
    int prev = -1;
    explicit DeviceGuard(int dev) { cudaGetDevice(&prev); if (prev != dev) cudaSetDevice(dev); else prev = -1; }
    ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

// One dense contraction of the network with its fused epilogue.
struct Gemm {
    int K = 0, N = 0;              // logical sizes
    int epi = cpb::EPI_ACT;
    std::vector<float> W;          // (K, N) row-major, host copy
    std::vector<float> p0, p1, p2; // epilogue constants (see build_gemms)
    // device, SIMT path
    float *dW = nullptr, *dp0 = nullptr, *dp1 = nullptr, *dp2 = nullptr;
    float* dWt = nullptr;          // fp32 path, tiny batches: the weights transposed, [N][Kp] with Kp = K rounded up to 4 (0 = none: K too deep)
    int Kp = 0;
    // device, UMMA path
    int k_pad = 0, k_stages = 0, nc = 0, n_chunks = 0;
    int max_nc = cpb::UM_MAX_NC;   // widest chunk the schedule wants for this contraction (256 = two TMEM slots, 128 = one)
    int model = 0;                 // emulator of a multi-emulator engine this contraction belongs to
    bool layer0 = true;            // first contraction of its emulator (operand = standardised parameters)
    bool last = true;              // last contraction of its emulator (writes the spectrum)
    float acc_scale = 1.f;         // 1 / (weight scale * input-activation scale), a power of two
    bool tri = false;              // upper-triangular weights (quadratic form): chunk c skips the K stages past its columns
    __half* dwpack = nullptr;
    float4* depi = nullptr;
    float4* depi_b = nullptr;      // binned-output constants (cpb200_engine_set_binning)
};

constexpr int64_t HOST_CHUNK_BYTES = 128ll << 20;  // output bytes per pipelined host chunk

}  // namespace

struct cpb200_engine {
    int device = 0, precision = 0;
    int P = 0, n_modes = 0, n_pcas = 0, num_sms = 0;      // of emulator 0 (the only one of an ordinary engine)
    std::vector<Gemm> gemms;
    float *d_pmean = nullptr, *d_pstd = nullptr;
    // multi-emulator engines (cpb200_create_multi): several emulators evaluated on the same rows by one launch
    int n_models = 1;
    int mP[cpb::UM_MAX_MODELS] = {0, 0, 0}, mModes[cpb::UM_MAX_MODELS] = {0, 0, 0};
    float* md_pmean[cpb::UM_MAX_MODELS] = {nullptr, nullptr, nullptr};
    float* md_pstd[cpb::UM_MAX_MODELS] = {nullptr, nullptr, nullptr};
    // SIMT scratch (two ping-pong activation buffers)
    float* simt_buf[2] = {nullptr, nullptr};
    int64_t simt_rows = 0;
    int simt_width = 0;
    // UMMA scratch
    uint8_t* um_scratch = nullptr;
    unsigned long long um_scratch_cta = 0;
    unsigned um_a0_bytes = 0, um_abuf_bytes = 0;
    // status words in mapped pinned host memory (readable even after a device trap has killed the context):
    // [0] watchdog role/barrier code, [1] range flag (a split operand left the fp16 range)
    int* h_status = nullptr;
    int* d_status = nullptr;                      // device alias of h_status
    unsigned long long* d_max_cycles = nullptr;   // fused kernel: SM cycles of the slowest CTA of the latest launch (timing on)
    // launches of one engine share its activation scratch: a launch on another stream than the previous one waits for it
    cudaEvent_t ev_order = nullptr;
    cudaStream_t last_stream = nullptr;
    bool last_valid = false;
    // quadratic-form use of the fused kernel (Planck-lite): set by the caller before launch_umma
    const uint8_t* qf_ext_a0 = nullptr;
    const float* qf_d = nullptr;
    long long qf_pitch = 0;
    int qf_cols = 0;
    float* qf_partial = nullptr;
    std::vector<unsigned> jobs;         // job table of the fused kernel (build_jobs), passed in the kernel parameters
    std::vector<unsigned short> jobs_x;
    int n_jobs = 0;
    int um_rot_mul = 0;                 // see UmmaParams::rot_mul
    unsigned um_pbuf_bytes = 0;         // see UmmaParams::pbuf_bytes
    int schedule = 0;                   // SCHED_* chosen by plan_schedule
    long long* d_counters = nullptr;    // debug counters, UM_CNT_SLOTS per CTA (allocated on demand)
    // host pipeline
    cudaStream_t hstream[2] = {nullptr, nullptr};
    cudaEvent_t ev_piece[16] = {};       // mid-size pageable results: one event per piece of the device->host transfer
    float* h_in[2] = {nullptr, nullptr};
    float* h_out[2] = {nullptr, nullptr};
    float* d_in[2] = {nullptr, nullptr};
    float* d_out[2] = {nullptr, nullptr};
    int64_t chunk_rows = 0;
    int64_t dpitch = 0;                 // row pitch of d_out (n_modes rounded up to 8 floats: 32-byte rows)
    bool pipeline_ready = false, bounce_ready = false;
    // bookkeeping
    int64_t launches = 0;
    bool timing = false;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    bool ev_valid = false;
};

namespace {

template <class T>
int upload(T** dptr, const T* h, size_t n) {
    CU(cudaMalloc(reinterpret_cast<void**>(dptr), std::max<size_t>(n, 1) * sizeof(T)));
    if (n) CU(cudaMemcpy(*dptr, h, n * sizeof(T), cudaMemcpyHostToDevice));
    return 0;
}

// Turn the reference attributes into the list of contractions + folded epilogue constants.
int build_gemms(const cpb200_model_desc* d, std::vector<Gemm>& out, int model = 0) {
    const int nd = d->n_dense;
    const bool pca = d->n_pcas > 0;
    const size_t first_gemm = out.size();
    for (int i = 0; i < nd; ++i) {
        Gemm g;
        g.K = d->dims[i];
        g.N = d->dims[i + 1];
        g.W.assign(d->W[i], d->W[i] + static_cast<size_t>(g.K) * g.N);
        const float* b = d->b[i];
        g.p0.resize(g.N); g.p1.resize(g.N); g.p2.assign(g.N, 0.f);
        if (i < nd - 1) {                       // hidden layer: NN.py:375-378
            g.epi = cpb::EPI_ACT;
            for (int n = 0; n < g.N; ++n) { g.p0[n] = b[n]; g.p1[n] = d->alphas[i][n]; g.p2[n] = d->betas[i][n]; }
        } else if (pca) {                       // PCA coefficients: (acc + b) * pca_std + pca_mean
            g.epi = cpb::EPI_AFFINE;
            for (int n = 0; n < g.N; ++n) {
                g.p0[n] = d->pca_std[n];
                g.p1[n] = static_cast<float>(static_cast<double>(b[n]) * d->pca_std[n] + d->pca_mean[n]);
            }
        } else {                                // NN output: (acc + b) * features_std + features_mean
            g.epi = cpb::EPI_OUT;
            for (int n = 0; n < g.N; ++n) {
                g.p0[n] = d->features_std[n];
                g.p1[n] = static_cast<float>(static_cast<double>(b[n]) * d->features_std[n] + d->features_mean[n]);
            }
        }
        out.push_back(std::move(g));
    }
    if (pca) {                                  // basis contraction: PCAplusNN.py:381
        Gemm g;
        g.K = d->n_pcas;
        g.N = d->n_modes;
        g.epi = cpb::EPI_OUT;
        g.W.assign(d->pca_transform_matrix, d->pca_transform_matrix + static_cast<size_t>(g.K) * g.N);
        g.p0.assign(d->features_std, d->features_std + g.N);
        g.p1.assign(d->features_mean, d->features_mean + g.N);
        g.p2.assign(g.N, 0.f);
        out.push_back(std::move(g));
    }
    for (size_t i = first_gemm; i < out.size(); ++i) { out[i].model = model; out[i].layer0 = i == first_gemm; out[i].last = i + 1 == out.size(); }
    return 0;
}

int round_up(int x, int m) { return (x + m - 1) / m * m; }

// Pack one contraction for the tensor-core path: fp16 hi/lo split, chunked along N, staged along
// K, each (chunk, stage, hi|lo) tile stored in the UMMA K-major core-matrix order
//   byte offset of element (n, k) inside a tile = (n/8)*512 + (k/8)*128 + (n%8)*16 + (k%8)*2.
//
// Range conditioning: fp16 keeps 11 significand bits only down to 6.1e-5, and the lo parts are
// 2^-11 of the value, so operands are pre-scaled by powers of two (exact) to sit high in the fp16
// range: weights by s_w (largest |w| lands in [16384, 32768)), the incoming activations by s_in,
// the produced activations by s_out.  The epilogue undoes s_w*s_in on the fp32 accumulator.
// chunking of one contraction: K padded to whole stages, N cut into equal chunks of at most max_nc columns (multiples of 32)
void plan_chunks(Gemm& g, int k_pad) {
    g.k_pad = k_pad;
    g.k_stages = k_pad / cpb::UM_KS;
    g.n_chunks = (g.N + g.max_nc - 1) / g.max_nc;
    g.nc = round_up((g.N + g.n_chunks - 1) / g.n_chunks, 32);
}

int pack_umma(Gemm& g, int k_pad, float s_in, float s_out) {
    float wmax = 0.f;
    for (float w : g.W) {
        if (!std::isfinite(w)) return fail(CPB200_E_RANGE, "non-finite weight");
        wmax = std::max(wmax, std::fabs(w));
    }
    int ex = wmax > 0.f ? static_cast<int>(std::floor(std::log2(32768.0 / wmax))) : 0;
    ex = std::max(-10, std::min(14, ex));
    const float s_w = std::ldexp(1.f, ex);
    if (wmax * s_w > 60000.f) return fail(CPB200_E_RANGE, "weight %g outside the fp16 split range", static_cast<double>(wmax));
    if (wmax > 0.f && wmax * s_w < 1024.f)      // the scale is capped at 2^14: a tinier matrix would lose its lo parts to fp16 underflow
        return fail(CPB200_E_RANGE, "largest weight %g of a layer is below the fp16 split range; rescale the layer "
                    "(fold the factor into features_std_ / the next layer)", static_cast<double>(wmax));
    // The tensor core truncates (does not round) the fp32 accumulator after each of the 3*K/16 MMA steps,
    // which shrinks |acc| by about half an ulp per step; (1 + bias_ulps * 2^-24) undoes the mean of that.
    // (the constant is pinned by tests/test_parity_gpu.py::test_truncation_bias_calibration)
    float bias_ulps = cpb::UM_TRUNC_BIAS_ULPS_PER_STEP * 3.f * static_cast<float>(k_pad / 16);
#ifdef CPB200_PROFILING_BUILD
    if (const char* envb = getenv("CPB200_BIAS_ULPS_PER_STEP")) bias_ulps = static_cast<float>(atof(envb)) * 3.f * static_cast<float>(k_pad / 16);
#endif
    const float inv = (1.f / (s_w * s_in)) * (1.f + bias_ulps * 5.9604645e-8f);
    g.acc_scale = inv;
    plan_chunks(g, k_pad);
    // per (chunk, stage): [CTA0: hi half | lo half][CTA1: hi half | lo half], a half = nc/2 columns
    const size_t tile = static_cast<size_t>(g.nc / 2) * cpb::UM_KS;        // elements per hi or lo half tile
    std::vector<__half> pack(static_cast<size_t>(g.n_chunks) * g.k_stages * 4 * tile, __float2half(0.f));
    for (int k = 0; k < g.K; ++k) {
        for (int n = 0; n < g.N; ++n) {
            const float w = g.W[static_cast<size_t>(k) * g.N + n] * s_w;
            const __half hi = __float2half_rn(w);
            const __half lo = __float2half_rn(w - __half2float(hi));
            // accumulator column order inside each 32-column block: position 8i+2t+cc holds logical
            // column 8t+2i+cc, which hands an epilogue lane 8 consecutive logical columns
            // (tcgen05.ld.16x256b fragment); the map is its own inverse
            const int c = n / g.nc, nl = n % g.nc, s = k / cpb::UM_KS, kk = k % cpb::UM_KS;
            const int nb = nl & 31;
            const int nn = (nl & ~31) | (((nb >> 1) & 3) << 3) | ((nb >> 3) << 1) | (nb & 1);
            const int half = g.nc / 2, r = nn / half, nh = nn % half;          // which CTA of the pair holds this column
            const size_t base = (static_cast<size_t>(c) * g.k_stages + s) * 4 * tile + static_cast<size_t>(r) * 2 * tile;
            const size_t off = static_cast<size_t>(nh / 8) * 256 + (kk / 8) * 64 + (nh % 8) * 8 + (kk % 8);
            pack[base + off] = hi;
            pack[base + tile + off] = lo;
        }
    }
    // Epilogue constants, stored in the order the epilogue lanes read them: inside each 32-column block lane
    // group t (= lane & 3) owns the 8 logical columns 8t..8t+7 as four column pairs i (columns 8t+2i, 8t+2i+1), and
    // every constant is stored per pair so that the epilogue's packed fp32 instructions (FFMA2 ...) find both
    // columns' values in one aligned register pair.  Operand-producing layers: two float4 per pair at positions
    // 4(2i) + t and 4(2i+1) + t of the block's 32 - ACT {b0, b1, a0, a1} and {c0, c1, d0, d1} (a = -alpha*log2e,
    // c = beta*s, d = (1-beta)*s), AFFINE {scale0, scale1, shift0, shift1} in the first.  Output layers: one float4
    // {scale0, scale1, shift0, shift1} per pair at position 4i + t of the block's 16; the whole array once for
    // predictions and once more multiplied by log2(10) for 10**y = 2**(y*log2 10).
    std::vector<float4> epi;
    auto put = [](float4& q, int cc, float lo, float hi) { if (cc) { q.y = lo; q.w = hi; } else { q.x = lo; q.z = hi; } };
    if (g.epi == cpb::EPI_OUT) {
        const size_t per_variant = static_cast<size_t>(g.n_chunks) * (g.nc / 2);
        epi.assign(2 * per_variant, make_float4(0.f, 0.f, 0.f, 0.f));
        const double l2_10 = 3.321928094887362;
        for (int n = 0; n < g.N; ++n) {
            const int nb = n & 31, t = nb >> 3, m = (nb & 7) >> 1;
            const size_t idx = static_cast<size_t>(n >> 5) * 16 + 4 * m + t;
            const float sc = g.p0[n] * inv, sh = g.p1[n];
            const float sc10 = static_cast<float>(static_cast<double>(g.p0[n]) * inv * l2_10);
            const float sh10 = static_cast<float>(static_cast<double>(g.p1[n]) * l2_10);
            put(epi[idx], n & 1, sc, sh);
            put(epi[per_variant + idx], n & 1, sc10, sh10);
        }
    } else {
        epi.assign(static_cast<size_t>(g.n_chunks) * g.nc, make_float4(0.f, 0.f, 0.f, 0.f));
        for (int n = 0; n < g.N; ++n) {
            const int nb = n & 31, t = nb >> 3, i = (nb & 7) >> 1, cc = nb & 1;
            float4& qa = epi[(n & ~31) + 4 * (2 * i) + t];
            float4& qb = epi[(n & ~31) + 4 * (2 * i + 1) + t];
            if (g.epi == cpb::EPI_ACT) {
                const float alpha = g.p1[n], beta = g.p2[n];
                // a = acc*acc_scale + b;  h*s_out = ((1-beta)*s_out * sigmoid + beta*s_out) * a
                put(qa, cc, g.p0[n], static_cast<float>(-static_cast<double>(alpha) * 1.4426950408889634));
                put(qb, cc, beta * s_out, (1.f - beta) * s_out);
            } else if (g.epi == cpb::EPI_AFFINE) {
                put(qa, cc, g.p0[n] * inv * s_out, g.p1[n] * s_out);
            }                                   // EPI_QUADFORM: no per-column constants
        }
    }
    if (int rc = upload(&g.dwpack, pack.data(), pack.size())) return rc;
    if (int rc = upload(&g.depi, epi.data(), epi.size())) return rc;
    return 0;
}

// ---------------------------------------------------------------------------------------------
// Job table of the fused kernel (UmmaParams::jobs): the static schedule of one iteration.
//
// A job = (contraction g, tile t of the CTA's pair, N-chunk c) = one accumulator in TMEM (one 128-column slot, or two
// for chunks wider than 128) drained by one epilogue.  The MMA issuer and the loader walk the table's entries in order;
// an entry is a PART of a job (K stages [s0, s0 + ns)), so a short job can be issued between two parts of a long one.
// The epilogue warps take the jobs in the order of their last parts.  Why: with two full-width accumulators a job whose
// MMA is short (layer 0: K = 32; cmb_PP's output chunks: K = 64) makes the job after it wait for the previous epilogue.
// Issued in the MIDDLE of a long job - as soon as the other accumulator has been drained - its own epilogue starts
// that much earlier and the job after the host finds its accumulator free.
//
// Schedules (plan_schedule):
//  SCHED_SIMPLE     round 1's order: whole jobs; cur = contractions 1 .. G-1 of pair it-1, next = layer 0 of pair it, spread
//                   over the output phase.
//  SCHED_MERGE_L0   same split, but each layer-0 job is issued inside an output-phase job of its tile (big-K output phase:
//                   cmb_TT/EE/TE, P(k)).
//  SCHED_OVERLAP    the last contraction has a short K (cmb_PP: 64): its chunks are store-bound, almost MMA-free.
//                   cur = the output chunks of pair it-1, next = the WHOLE trunk of pair it; one output job is issued inside
//                   every two-slot trunk job, the rest surround the layer-0 and coefficient jobs.  The output phase's
//                   operand (written by pair p's trunk, read during pair p+1's) lives in per-parity buffers.
// ---------------------------------------------------------------------------------------------
enum { SCHED_SIMPLE = 0, SCHED_MERGE_L0 = 1, SCHED_OVERLAP = 2 };

struct JobEntry {
    int g = 0, t = 0, c = 0, nxt = 0, s0 = 0, ns = 0, first = 1, last = 1, slot = 0, two = 0, in_code = 0, out_code = 0;
    int job = -1;                   // index of the job this entry is a part of (build-time only)
};

// which schedule fits the model; also fixes the chunk width of every contraction (before pack_umma)
int plan_schedule(std::vector<Gemm>& gemms, int k_pad0) {
    const int G = static_cast<int>(gemms.size());
    int sched = SCHED_SIMPLE;
#ifdef CPB200_PROFILING_BUILD
    const char* force = getenv("CPB200_SCHEDULE");      // profiling build only: "simple", "merge", "overlap"
#else
    const char* force = nullptr;
#endif
    (void)k_pad0;
    // Measured on B200 (profiles/r2c_schedule_ab.txt, SM cycles of a 1 048 576-row launch, SIMPLE / MERGE_L0 / OVERLAP):
    // cmb_TT 11.81 / 11.66 / -, cmb_TE 13.27 / 13.06 / -, PKLIN 4.885 / 4.817 / -, cmb_PP 9.12 / 9.06 / 9.03-9.11 M.
    // MERGE_L0 is the default wherever there is an output phase to host layer 0; OVERLAP buys cmb_PP nothing measurable
    // (the job after a hosted output job still waits for that job's epilogue: two accumulators) and stays a
    // profiling-build option.
    if (G >= 2 && gemms[0].epi != cpb::EPI_QUADFORM) sched = SCHED_MERGE_L0;
    if (gemms.back().model > 0) force = nullptr;          // several emulators in one launch: MERGE_L0 is the only order built for it
    if (force) {
        if (!strcmp(force, "simple")) sched = SCHED_SIMPLE;
        else if (!strcmp(force, "merge") && G >= 2) sched = SCHED_MERGE_L0;
        else if (!strcmp(force, "overlap") && G >= 3) sched = SCHED_OVERLAP;
    }
    for (Gemm& g : gemms) g.max_nc = cpb::UM_MAX_NC;
    return sched;
}

// measured costs (SM cycles) the placement of inserted jobs is derived from
double est_epilogue_cycles(const Gemm& L) {
    const double per256 = L.epi == cpb::EPI_ACT ? 8500.0 : L.epi == cpb::EPI_AFFINE ? 3000.0 : 5300.0;
    return 1000.0 + per256 * L.nc / 256.0;
}
double est_stage_cycles(const Gemm& L) { return 6.0 * std::max(L.nc / 2.0, 32.0); }       // 6 MMAs of nc/2 cycles; A-operand read floor below N = 64

// cut a job (entry with all its stages) into parts at the given stage indices
std::vector<JobEntry> split_job(const JobEntry& e, const std::vector<int>& cuts) {
    std::vector<int> bounds;
    for (int c : cuts) if (c > 0 && c < e.ns && (bounds.empty() || c > bounds.back())) bounds.push_back(c);
    bounds.push_back(e.ns);
    std::vector<JobEntry> parts;
    int s = 0;
    for (size_t i = 0; i < bounds.size(); ++i) {
        JobEntry p = e;
        p.s0 = e.s0 + s; p.ns = bounds[i] - s; p.first = i == 0; p.last = i + 1 == bounds.size();
        parts.push_back(p);
        s = bounds[i];
    }
    return parts;
}

// `host` with `guest` issued inside it: after the K stage at which the epilogue of `prev` (the job before the host, whose
// accumulator the guest takes over) has finished.  When that is beyond the host's last stage the guest simply follows.
void push_hosted(std::vector<JobEntry>& out, const std::vector<Gemm>& gemms, const JobEntry& host, const JobEntry& guest,
                 const JobEntry* prev) {
    const double wait = prev ? est_epilogue_cycles(gemms[prev->g]) : 0.0;
    const int cut = static_cast<int>(wait / est_stage_cycles(gemms[host.g])) + 1;
    if (cut >= host.ns) { out.push_back(host); out.push_back(guest); return; }
    std::vector<JobEntry> parts = split_job(host, {cut});
    out.push_back(parts[0]);
    out.push_back(guest);
    out.push_back(parts[1]);
}

// TMEM slots for every job of an ordered entry list.  A job owns its slot(s) from its first part to the end of its
// epilogue; the kernel's barriers make ANY static assignment correct as long as a job never waits for a slot whose
// owner has not been issued completely (deadlock), i.e. the previous owner's last part must precede the new job's first
// part.  Among the candidates the allocator takes the slot whose previous owner completes earliest in the epilogue order
// (ties: the lower slot).
int assign_slots(std::vector<JobEntry>& entries) {
    const int n = static_cast<int>(entries.size());
    int owner_last[cpb::UM_TMEM_SLOTS];          // table position of the last part of the slot's current owner
    bool open[cpb::UM_TMEM_SLOTS];               // the owner's last part has not been seen yet
    // the table repeats every iteration: run the allocator twice and keep the second pass (steady state), with the
    // ownership at the end of pass one (positions shifted by -n) as the initial condition
    std::vector<int> job_slot;
    for (int sl = 0; sl < cpb::UM_TMEM_SLOTS; ++sl) { owner_last[sl] = -2 * n - (cpb::UM_TMEM_SLOTS - sl); open[sl] = false; }
    for (int pass = 0; pass < 2; ++pass) {
        job_slot.assign(n, -1);
        for (int i = 0; i < n; ++i) {
            JobEntry& e = entries[i];
            if (e.first) {
                int last_pos = i;
                while (!(entries[last_pos].last && entries[last_pos].job == e.job)) ++last_pos;
                int best = -1, best_key = 0;
                if (e.two) {
                    for (int sl = 0; sl < cpb::UM_TMEM_SLOTS; sl += 2) {
                        if (open[sl] || open[sl + 1]) continue;
                        const int key = std::max(owner_last[sl], owner_last[sl + 1]);
                        if (best < 0 || key < best_key) { best = sl; best_key = key; }
                    }
                } else {
                    for (int sl = 0; sl < cpb::UM_TMEM_SLOTS; ++sl) {
                        if (open[sl]) continue;
                        if (best < 0 || owner_last[sl] < best_key) { best = sl; best_key = owner_last[sl]; }
                    }
                }
                if (best < 0) return fail(CPB200_E_UNSUPPORTED, "job table needs more than %d TMEM slots at entry %d", cpb::UM_TMEM_SLOTS, i);
                job_slot[e.job] = best;
                for (int k = 0; k <= e.two; ++k) { open[best + k] = true; owner_last[best + k] = last_pos + (pass ? 0 : -n); }
            }
            e.slot = job_slot[e.job];
            if (e.slot < 0) return fail(CPB200_E_UNSUPPORTED, "job table: part before its job's first part at entry %d", i);
            if (e.last) for (int k = 0; k <= e.two; ++k) open[e.slot + k] = false;
        }
    }
    return 0;
}

int build_jobs(const std::vector<Gemm>& gemms, int sched, std::vector<JobEntry>& out, int* rot_mul, int* par_stages) {
    const int G = static_cast<int>(gemms.size());
    const int n_models = gemms.back().model + 1;
    out.clear();
    *rot_mul = 0; *par_stages = 0;
    int n_jobs = 0;
    // per emulator: its range of contractions and the offset of its walk through the rotating buffers (an emulator with
    // G contractions advances the rotation by G - 1, see below)
    std::vector<int> g0(n_models + 1, G), base(n_models + 1, 0);
    for (int g = G - 1; g >= 0; --g) g0[gemms[g].model] = g;
    for (int m = 0; m < n_models; ++m) base[m + 1] = base[m] + (g0[m + 1] - g0[m] - 1);
    auto make = [&](int g, int t, int c, int nxt) {
        JobEntry e;
        const Gemm& L = gemms[g];
        e.g = g; e.t = t; e.c = c; e.nxt = nxt; e.s0 = 0;
        e.ns = L.tri ? std::min(L.k_stages, ((c + 1) * L.nc) / cpb::UM_KS) : L.k_stages;      // upper-triangular weights: stages up to the chunk's last column
        e.two = L.nc > cpb::UM_SLOT_COLS;
        e.job = n_jobs++;
        // operand buffers: layer 0 reads the standardiser's tile of (emulator, tile); step n = 2(layer-1) + t of an emulator
        // reads rotating buffer 2n and writes 2n + 1, shifted by what the emulators before it (this pair) have advanced
        const int gl = g - g0[L.model], n = 2 * (gl - 1) + t, sh = base[L.model];
        e.in_code = L.layer0 ? cpb::UM_BUF_A0 + 2 * L.model + t : cpb::UM_BUF_ROT + (((2 * n + sh) % 3) + 3) % 3;
        e.out_code = L.last ? cpb::UM_BUF_NONE : cpb::UM_BUF_ROT + (((2 * n + 1 + sh) % 3) + 3) % 3;
        if (sched == SCHED_OVERLAP) {
            if (g == G - 1) e.in_code = cpb::UM_BUF_PAR + t;
            if (g == G - 2) e.out_code = cpb::UM_BUF_PAR + t;
        }
        return e;
    };
    if (n_models > 1) {
        if (sched != SCHED_MERGE_L0) return fail(CPB200_E_UNSUPPORTED, "several emulators in one launch need the merged-layer-0 order");
        for (int m = 0; m < n_models; ++m)
            if (g0[m + 1] - g0[m] < 2) return fail(CPB200_E_UNSUPPORTED, "an emulator without a hidden layer cannot share a launch");
    }
    if (sched == SCHED_MERGE_L0) {
        // Per pair of tiles the emulators run one after the other: trunk and output phase of emulator m, and inside its
        // output jobs the layer-0 jobs of the emulator that follows - emulator m + 1 on the same pair, or emulator 0 on the
        // NEXT pair.  Every layer 0 thus runs just before its own trunk, in an output phase whose long MMAs hide its long
        // epilogue, and lands in the rotating buffer that is free then (tile 1's only after tile 0's last output chunk:
        // the buffer it writes is tile 0's output operand).
        *rot_mul = base[n_models] % 3;
        const JobEntry* prev = nullptr;
        std::vector<JobEntry> keep;                      // stable storage for `prev`
        keep.reserve(4096);
        for (int m = 0; m < n_models; ++m) {
            const int nm = (m + 1) % n_models, gl0 = g0[nm], glast = g0[m + 1] - 1;
            for (int g = g0[m] + 1; g < glast; ++g)
                for (int t = 0; t < 2; ++t)
                    for (int c = 0; c < gemms[g].n_chunks; ++c) { keep.push_back(make(g, t, c, 0)); out.push_back(keep.back()); prev = &keep.back(); }
            for (int t = 0; t < 2; ++t) {
                std::vector<JobEntry> outs, l0s;
                for (int c = 0; c < gemms[glast].n_chunks; ++c) outs.push_back(make(glast, t, c, 0));
                for (int c = 0; c < gemms[gl0].n_chunks; ++c) l0s.push_back(make(gl0, t, c, nm == 0 ? 1 : 0));
                const size_t n_o = outs.size(), n_l = l0s.size();
                std::vector<std::vector<JobEntry>> where(n_o);
                // spread over the tile's output jobs, leaving out the first one when there is room (it follows a long epilogue)
                const size_t lo = (n_o > n_l) ? 1 : 0;
                for (size_t k = 0; k < n_l; ++k) where[std::min(n_o - 1, lo + static_cast<size_t>((k + 0.5) * (n_o - lo) / n_l))].push_back(l0s[k]);
                for (size_t i = 0; i < n_o; ++i) {
                    keep.push_back(outs[i]);
                    const JobEntry& host = keep.back();
                    if (where[i].empty()) { out.push_back(host); prev = &host; continue; }
                    push_hosted(out, gemms, host, where[i][0], prev);
                    for (size_t k = 1; k < where[i].size(); ++k) out.push_back(where[i][k]);
                    prev = &host;
                }
            }
        }
    } else {
        std::vector<JobEntry> l0, head, tail;
        for (int g = 0; g < G; ++g)
            for (int t = 0; t < 2; ++t)
                for (int c = 0; c < gemms[g].n_chunks; ++c) {
                    if (sched == SCHED_OVERLAP) (g == G - 1 ? tail : head).push_back(make(g, t, c, g != G - 1));
                    else if (G == 1 || g == 0) l0.push_back(make(g, t, c, 1));
                    else (g == G - 1 ? tail : head).push_back(make(g, t, c, 0));
                }
        if (sched == SCHED_OVERLAP) {
            *par_stages = gemms[G - 1].k_stages;
            // hosts: the trunk jobs with a long MMA (at least 8 stages); every host takes one output job of the previous pair
            size_t n_host = 0;
            for (const JobEntry& e : head) n_host += e.ns >= 8;
            const size_t hosted = std::min(tail.size(), n_host);
            // what is left surrounds the other trunk jobs: before the first host (the layer-0 jobs: long epilogues, no MMA to
            // speak of) and after the last (the dependent coefficient jobs)
            const size_t rest = tail.size() - hosted;
            size_t n_pre = 0, n_post = 0;
            { bool seen_host = false; for (const JobEntry& e : head) { if (e.ns >= 8) seen_host = true; else (seen_host ? n_post : n_pre)++; } }
            size_t rest_post = n_post ? std::min(rest, std::max<size_t>(rest / 3, n_post)) : 0, rest_pre = n_pre ? rest - rest_post : 0;
            if (!n_pre) rest_post = n_post ? rest : 0;
            size_t oi = 0, pre_seen = 0, post_seen = 0, pre_used = 0, post_used = 0;
            const JobEntry* prev = nullptr;
            for (size_t i = 0; i < head.size(); ++i) {
                const JobEntry& e = head[i];
                if (e.ns >= 8 && oi < tail.size() && hosted > 0 && (oi - pre_used - post_used) < hosted) {
                    push_hosted(out, gemms, e, tail[oi++], prev);
                    prev = &head[i];
                    continue;
                }
                out.push_back(e);
                prev = &head[i];
                if (e.ns >= 8) continue;
                const bool is_pre = post_seen == 0 && pre_seen < n_pre;
                size_t want;
                if (is_pre) { ++pre_seen; want = rest_pre * pre_seen / n_pre - pre_used; }
                else { ++post_seen; want = rest_post * post_seen / std::max<size_t>(1, n_post) - post_used; }
                for (size_t k = 0; k < want && oi < tail.size(); ++k) { prev = &tail[oi]; out.push_back(tail[oi++]); (is_pre ? pre_used : post_used)++; }
            }
            while (oi < tail.size()) out.push_back(tail[oi++]);
        } else {
            // round-1 order: whole jobs; layer 0 of the next pair spread over the output phase
            *rot_mul = G > 1 ? (G - 1) % 3 : 0;
            out = head;
            size_t k = 0;
            for (size_t i = 0; i < tail.size(); ++i) {
                out.push_back(tail[i]);
                while (k < l0.size() && (k + 1) * tail.size() <= (i + 1) * (l0.size() + 1)) out.push_back(l0[k++]);
            }
            while (k < l0.size()) out.push_back(l0[k++]);
        }
    }
    if (out.size() > static_cast<size_t>(cpb::UM_MAX_JOBS))
        return fail(CPB200_E_UNSUPPORTED, "%zu job-table entries per tile pair exceed %d", out.size(), cpb::UM_MAX_JOBS);
    for (const JobEntry& e : out)
        if (e.c > 127 || e.s0 > 63 || e.ns > 63 || e.g > 15) return fail(CPB200_E_UNSUPPORTED, "layer too wide or too deep for the job-table encoding");
    return assign_slots(out);
}

void encode_jobs(const std::vector<JobEntry>& entries, std::vector<unsigned>& w0, std::vector<unsigned short>& w1) {
    w0.clear(); w1.clear();
    for (const JobEntry& e : entries) {
        w0.push_back(cpb::um_job_w0(e.g, e.t, e.c, e.nxt, e.s0, e.ns, e.first, e.last, e.slot, e.two));
        w1.push_back(cpb::um_job_w1(e.in_code, e.out_code));
    }
}

// Every launch of an engine uses the engine's activation scratch, and the entry points are asynchronous on the
// caller's stream: a launch on a different stream than the previous one is ordered behind it with an event (recorded
// after each launch; free when the stream does not change).  Inside a stream capture the event would belong to the
// graph, so captured launches neither record nor wait: a captured engine must not be used eagerly at the same time.
int order_before(cpb200_engine* e, cudaStream_t st) {
    if (e->last_valid && e->last_stream != st) {
        cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
        if (cudaStreamIsCapturing(st, &cs) == cudaSuccess && cs == cudaStreamCaptureStatusNone)
            CU(cudaStreamWaitEvent(st, e->ev_order, 0));
    }
    return 0;
}
int order_after(cpb200_engine* e, cudaStream_t st) {
    cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
    if (cudaStreamIsCapturing(st, &cs) != cudaSuccess || cs != cudaStreamCaptureStatusNone) { cudaGetLastError(); return 0; }
    if (!e->ev_order) CU(cudaEventCreateWithFlags(&e->ev_order, cudaEventDisableTiming));
    CU(cudaEventRecord(e->ev_order, st));
    e->last_stream = st;
    e->last_valid = true;
    return 0;
}

// what the kernel's watchdog code (UmmaParams::error_flag[0]) says: which role was stuck on which barrier
const char* watchdog_text(int code) {
    switch (code) {
        case 0: return "no watchdog code";
        case 101: return "loader waiting for the layer-0 operand (standardiser)";
        case 102: return "loader waiting for an activation tile (epilogue of the previous layer)";
        case 103: return "loader waiting for a free pipeline stage (MMA issuer)";
        case 104: return "loader waiting for a free epilogue-constant slot (epilogue)";
        case 201: return "MMA issuer waiting for a free accumulator (epilogue)";
        case 202: return "MMA issuer waiting for a loaded stage (loader / peer relay)";
        case 203: return "peer relay waiting for a loaded stage (loader)";
        case 301: return "standardiser waiting for its operand buffer (MMA issuer)";
        case 401: return "epilogue waiting for a finished accumulator (MMA issuer)";
        case 402: return "epilogue waiting for its constants (loader)";
        default: return "unknown role";
    }
}

// a failed synchronisation: the CUDA error plus, when the kernel's watchdog fired, the stuck role
int fail_sync(cpb200_engine* e, cudaError_t err, const char* what) {
    const int code = e && e->h_status ? *reinterpret_cast<volatile int*>(e->h_status) : 0;
    if (code) return fail(CPB200_E_CUDA, "%s failed: %s; fused-kernel watchdog code %d: %s", what, cudaGetErrorString(err), code, watchdog_text(code));
    return fail(CPB200_E_CUDA, "%s failed: %s", what, cudaGetErrorString(err));
}

// after a completed batch: was any split operand outside the fp16 range? (clears the flag)
int check_range(cpb200_engine* e) {
    if (!e->h_status) return 0;
    volatile int* f = reinterpret_cast<volatile int*>(e->h_status) + 1;
    if (*f == 0) return 0;
    *f = 0;
    return fail(CPB200_E_RANGE, "a parameter row is more than 64 standard deviations from the training mean (or not finite), or "
                "an activation left the fp16 split range: the tensor-core result of this batch is invalid; use CPB200_PREC_FP32_SIMT");
}

int launch_simt(cpb200_engine* e, const float* params, int64_t n, float* out, int64_t pitch, int flags,
                cudaStream_t st) {
    const int G = static_cast<int>(e->gemms.size());
    bool small = n <= cpb::SMALL_M;
    for (const Gemm& g : e->gemms) small = small && g.dWt != nullptr;
    if (small) {       // a sampler's single point: one warp per output column, the whole device on every layer (simt_kernels.cuh)
        const float* A = params;
        int64_t lda = e->P;
        const int m = static_cast<int>(n);
        for (int gi = 0; gi < G; ++gi) {
            const Gemm& g = e->gemms[gi];
            const bool last = gi == G - 1;
            float* C = last ? out : e->simt_buf[gi & 1];
            const int64_t ldc = last ? pitch : e->simt_width;
            const unsigned grid = static_cast<unsigned>((g.N + cpb::SMALL_WARPS - 1) / cpb::SMALL_WARPS);
            const float* im = gi == 0 ? e->d_pmean : nullptr;
            const float* is = gi == 0 ? e->d_pstd : nullptr;
            const int ten = (flags & CPB200_FLAG_TEN_TO) ? 1 : 0;
            // programmatic dependent launch: layer l + 1 may start (and fetch its weights) while layer l drains
            cudaLaunchConfig_t cfg = {};
            cfg.gridDim = dim3(grid); cfg.blockDim = dim3(32 * cpb::SMALL_WARPS); cfg.dynamicSmemBytes = 0; cfg.stream = st;
            cudaLaunchAttribute attr[1];
            attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
            attr[0].val.programmaticStreamSerializationAllowed = gi > 0 ? 1 : 0;
            cfg.attrs = attr; cfg.numAttrs = 1;
            const int acc_flag = (flags & CPB200_FLAG_ACCUMULATE) ? 1 : 0, zero = 0, m_i = m, K_i = g.K, Kp_i = g.Kp, N_i = g.N;
            const float* Wt = g.dWt; const float* q0 = g.dp0; const float* q1 = g.dp1; const float* q2 = g.dp2;
            cudaError_t lerr;
            if (g.epi == cpb::EPI_ACT)
                lerr = cudaLaunchKernelEx(&cfg, cpb::dense_small_kernel<cpb::EPI_ACT>, A, lda, Wt, K_i, Kp_i, N_i, C, ldc, m_i, q0, q1, q2, zero, zero, im, is);
            else if (g.epi == cpb::EPI_AFFINE)
                lerr = cudaLaunchKernelEx(&cfg, cpb::dense_small_kernel<cpb::EPI_AFFINE>, A, lda, Wt, K_i, Kp_i, N_i, C, ldc, m_i, q0, q1, q2, zero, zero, im, is);
            else
                lerr = cudaLaunchKernelEx(&cfg, cpb::dense_small_kernel<cpb::EPI_OUT>, A, lda, Wt, K_i, Kp_i, N_i, C, ldc, m_i, q0, q1, q2, ten, acc_flag, im, is);
            if (lerr != cudaSuccess) return fail(CPB200_E_CUDA, "tiny-batch layer launch failed: %s", cudaGetErrorString(lerr));
            ++e->launches;
            A = C;
            lda = ldc;
        }
        CU(cudaGetLastError());
        return 0;
    }
    for (int64_t r0 = 0; r0 < n; r0 += e->simt_rows) {
        const int64_t m = std::min<int64_t>(e->simt_rows, n - r0);
        const float* A = params + r0 * e->P;
        int64_t lda = e->P;
        for (int gi = 0; gi < G; ++gi) {
            const Gemm& g = e->gemms[gi];
            const bool last = gi == G - 1;
            float* C = last ? out + r0 * pitch : e->simt_buf[gi & 1];
            const int64_t ldc = last ? pitch : e->simt_width;
            dim3 grid((g.N + cpb::SIMT_BN - 1) / cpb::SIMT_BN, static_cast<unsigned>((m + cpb::SIMT_BM - 1) / cpb::SIMT_BM));
            const float* im = gi == 0 ? e->d_pmean : nullptr;
            const float* is = gi == 0 ? e->d_pstd : nullptr;
            const int ten = (flags & CPB200_FLAG_TEN_TO) ? 1 : 0;
            if (g.epi == cpb::EPI_ACT)
                cpb::dense_simt_kernel<cpb::EPI_ACT><<<grid, 256, 0, st>>>(A, lda, g.dW, g.K, g.N, C, ldc, m, g.dp0, g.dp1, g.dp2, 0, 0, im, is);
            else if (g.epi == cpb::EPI_AFFINE)
                cpb::dense_simt_kernel<cpb::EPI_AFFINE><<<grid, 256, 0, st>>>(A, lda, g.dW, g.K, g.N, C, ldc, m, g.dp0, g.dp1, g.dp2, 0, 0, im, is);
            else
                cpb::dense_simt_kernel<cpb::EPI_OUT><<<grid, 256, 0, st>>>(A, lda, g.dW, g.K, g.N, C, ldc, m, g.dp0, g.dp1, g.dp2, ten, (flags & CPB200_FLAG_ACCUMULATE) ? 1 : 0, im, is);
            ++e->launches;
            A = C;
            lda = ldc;
        }
    }
    CU(cudaGetLastError());
    return 0;
}

int launch_umma(cpb200_engine* e, const float* const* params, int64_t n, float* const* out, const int64_t* pitch, const int* flags,
                cudaStream_t st) {
    cpb::UmmaParams p;
    memset(&p, 0, sizeof p);
    p.n_gemms = static_cast<int>(e->gemms.size());
    p.n_models = e->n_models;
    for (int m = 0; m < e->n_models; ++m) {
        cpb::UmmaModel& M = p.m[m];
        M.params = params ? params[m] : nullptr;
        M.pmean = e->md_pmean[m];
        M.pstd = e->md_pstd[m];
        M.out = out ? out[m] : nullptr;
        M.out_pitch = pitch ? pitch[m] : 0;
        M.n_parameters = e->mP[m];
        M.n_modes = e->mModes[m];
        M.ten_to = (flags && (flags[m] & CPB200_FLAG_TEN_TO)) ? 1 : 0;
        M.accumulate = (flags && (flags[m] & CPB200_FLAG_ACCUMULATE)) ? 1 : 0;
        M.binned = (flags && (flags[m] & CPB200_FLAG_BINNED)) ? 1 : 0;
    }
    for (int i = 0; i < p.n_gemms; ++i) {
        const Gemm& g = e->gemms[i];
        p.g[i].wpack = g.dwpack;
        p.g[i].epi = g.depi;
        p.g[i].epi_b = g.depi_b;
        p.g[i].k_stages = g.k_stages;
        p.g[i].nc = g.nc;
        p.g[i].n_chunks = g.n_chunks;
        p.g[i].epi_kind = g.epi;
        p.g[i].acc_scale = g.acc_scale;
        p.g[i].tri = g.tri ? 1 : 0;
        p.g[i].model = static_cast<short>(g.model);
        p.g[i].layer0 = g.layer0 ? 1 : 0;
    }
    p.debug_flags = 0;
    p.discard_ok = 1;
    for (size_t i = 0; i < e->gemms.size(); ++i) if (!e->gemms[i].last && e->gemms[i].nc % 64) p.discard_ok = 0;
    if (e->d_counters) if (const char* dbg = getenv("CPB200_DEBUG_EPILOGUE")) p.debug_flags = atoi(dbg);   // profiling ablations
    p.n_rows = n;
    const int64_t n_tiles = (n + cpb::UM_TILE_M - 1) / cpb::UM_TILE_M;
    p.n_pairs = (n_tiles + 1) / 2;
    p.n_pairgroups = (p.n_pairs + cpb::UM_CLUSTER - 1) / cpb::UM_CLUSTER;
    p.scratch = e->um_scratch;
    p.scratch_cta_bytes = e->um_scratch_cta;
    p.a0_bytes = e->um_a0_bytes;
    p.abuf_bytes = e->um_abuf_bytes;
    for (int i = 0; i < e->n_jobs; ++i) { p.jobs[i] = e->jobs[i]; p.jobs_x[i] = e->jobs_x[i]; }
    p.n_jobs = e->n_jobs;
    p.rot_mul = e->um_rot_mul;
    p.pbuf_bytes = e->um_pbuf_bytes;
    p.error_flag = e->d_status;
    p.max_cycles = e->timing ? e->d_max_cycles : nullptr;       // the cycle figure costs a memset + one atomic per CTA: timing mode only
    p.ext_a0 = e->qf_ext_a0; p.qf_d = e->qf_d; p.qf_pitch = e->qf_pitch; p.qf_cols = e->qf_cols; p.qf_partial = e->qf_partial;
    if (p.max_cycles) cudaMemsetAsync(e->d_max_cycles, 0, sizeof(unsigned long long), st);
    p.counters = e->d_counters;
#ifdef CPB200_PROFILING_BUILD
    if (e->d_counters) if (const char* al = getenv("CPB200_ALIGN_EVERY")) {      // phase-alignment probe (see UmmaParams::align_every)
        static unsigned* d_align = nullptr;
        if (!d_align) cudaMalloc(reinterpret_cast<void**>(&d_align), sizeof(unsigned));
        if (d_align) { cudaMemsetAsync(d_align, 0, sizeof(unsigned), st); p.align_every = atoi(al); p.align_counter = d_align; }
    }
#endif
    int max_clusters = e->num_sms / cpb::UM_CLUSTER;
    if (const char* cap = getenv("CPB200_DEBUG_MAX_CLUSTERS")) max_clusters = std::max(1, std::min(max_clusters, atoi(cap)));  // tuning aid
    // Small batches leave most SMs idle: deal the output contraction's chunks (57 % of a cmb_TT pair's cycles) out to
    // up to n_chunks clusters per pair-group; each of them walks the trunk on its own copy of the activations.
    p.out_split = 1;
    {
        const Gemm& last = e->gemms.back();
        const int64_t room = max_clusters / std::max<int64_t>(1, p.n_pairgroups);
        if (e->n_models == 1 && e->gemms.size() > 1 && last.epi == cpb::EPI_OUT && last.n_chunks > 1 && room >= 2 && !getenv("CPB200_NO_OUT_SPLIT")) {
            const int smax = static_cast<int>(std::min<int64_t>(room, last.n_chunks));
            const int per = (last.n_chunks + smax - 1) / smax;            // chunks per cluster at the widest split
            p.out_split = (last.n_chunks + per - 1) / per;                // the narrowest split that reaches it
        }
    }
    const int grid = static_cast<int>(std::min<int64_t>(p.n_pairgroups, max_clusters)) * p.out_split * cpb::UM_CLUSTER;
    int mode = cpb::UM_MODE_PLAIN;
    for (int m = 0; m < e->n_models; ++m) {
        if (p.m[m].binned) mode = cpb::UM_MODE_BINNED;
        else if (p.m[m].accumulate && mode == cpb::UM_MODE_PLAIN) mode = cpb::UM_MODE_ACC;
    }
    const int variant = (p.counters ? 6 : 0) + (p.out_split > 1 ? 3 : 0) + mode;
#define UM_LAUNCH(PROF, SPLIT, MODE) cpb::fused_mlp_umma_kernel<PROF, SPLIT, MODE><<<grid, cpb::UM_THREADS, cpb::UM_SMEM_BYTES, st>>>(p)
    switch (variant) {
        case 0: UM_LAUNCH(false, false, 0); break;
        case 1: UM_LAUNCH(false, false, 1); break;
        case 2: UM_LAUNCH(false, false, 2); break;
        case 3: UM_LAUNCH(false, true, 0); break;
        case 4: UM_LAUNCH(false, true, 1); break;
        case 5: UM_LAUNCH(false, true, 2); break;
        case 6: UM_LAUNCH(true, false, 0); break;
        case 7: UM_LAUNCH(true, false, 1); break;
        case 8: UM_LAUNCH(true, false, 2); break;
        case 9: UM_LAUNCH(true, true, 0); break;
        case 10: UM_LAUNCH(true, true, 1); break;
        default: UM_LAUNCH(true, true, 2); break;
    }
#undef UM_LAUNCH
    ++e->launches;
    CU(cudaGetLastError());
    return 0;
}

int forward_async(cpb200_engine* e, const float* params, int64_t n, float* out, int64_t pitch, int flags,
                  cudaStream_t st) {
    if (n == 0) return 0;
    if (int rc = order_before(e, st)) return rc;
    const int64_t pitch64 = pitch;
    if (int rc = e->precision == CPB200_PREC_FP32_SIMT ? launch_simt(e, params, n, out, pitch, flags, st)
                                                       : launch_umma(e, &params, n, &out, &pitch64, &flags, st)) return rc;
    return order_after(e, st);
}

}  // namespace

namespace {
// Tensor-core path set-up shared by the emulator engines and the likelihood's quadratic form: pack every
// contraction, size the per-CTA scratch, build the job table, opt in to the kernel's shared memory.
int setup_umma(cpb200_engine* e) {
    int rc;
    int k_pad = round_up(e->P, cpb::UM_KS);
    unsigned max_stages = 1, a0_stages = 1;
    e->schedule = plan_schedule(e->gemms, k_pad);
    for (size_t i = 0; i < e->gemms.size(); ++i) {
        Gemm& g = e->gemms[i];
        if (g.layer0) k_pad = round_up(g.K, cpb::UM_KS);       // an emulator's first contraction: K = its parameter count
        const float s_in = g.layer0 ? cpb::UM_INPUT_SCALE : cpb::UM_ACT_SCALE;
        if ((rc = pack_umma(g, k_pad, s_in, cpb::UM_ACT_SCALE))) return rc;
        if (g.layer0) a0_stages = std::max<unsigned>(a0_stages, g.k_stages);
        else max_stages = std::max<unsigned>(max_stages, g.k_stages);
        k_pad = g.n_chunks * g.nc;          // the next contraction's (padded) K
    }
    e->um_a0_bytes = a0_stages * cpb::UM_A_STAGE_BYTES;
    e->um_abuf_bytes = max_stages * cpb::UM_A_STAGE_BYTES;
    {
        std::vector<JobEntry> entries;
        int par_stages = 0;
        if ((rc = build_jobs(e->gemms, e->schedule, entries, &e->um_rot_mul, &par_stages))) return rc;
        encode_jobs(entries, e->jobs, e->jobs_x);
        e->n_jobs = static_cast<int>(entries.size());
        e->um_pbuf_bytes = static_cast<unsigned>(par_stages) * cpb::UM_A_STAGE_BYTES;
        if (e->schedule == SCHED_OVERLAP) {
            // the last contraction's operand lives in the parity buffers: the rotating buffers only hold the trunk's tiles
            max_stages = 1;
            for (size_t i = 1; i + 1 < e->gemms.size(); ++i) max_stages = std::max<unsigned>(max_stages, e->gemms[i].k_stages);
            e->um_abuf_bytes = max_stages * cpb::UM_A_STAGE_BYTES;
        }
    }
    e->um_scratch_cta = 2ull * e->n_models * e->um_a0_bytes + static_cast<unsigned long long>(cpb::UM_ACT_BUFS) * e->um_abuf_bytes + 4ull * e->um_pbuf_bytes;
#ifdef CPB200_PROFILING_BUILD
    if (const char* pad = getenv("CPB200_SCRATCH_PAD")) e->um_scratch_cta += static_cast<unsigned long long>(atoll(pad)) / 128 * 128;   // L2 set-aliasing probe
#endif
    if (e->gemms.size() == 1 && e->gemms[0].epi == cpb::EPI_QUADFORM) e->um_scratch_cta = 1024;   // operand tiles are prebuilt elsewhere
    const size_t total = static_cast<size_t>(e->um_scratch_cta) * e->num_sms;
    if (cudaMalloc(reinterpret_cast<void**>(&e->um_scratch), total) != cudaSuccess)
        return (fail(CPB200_E_CUDA, "scratch allocation failed: %s", cudaGetErrorString(cudaGetLastError())));
    if (cudaMemset(e->um_scratch, 0, total) != cudaSuccess) return (fail(CPB200_E_CUDA, "memset failed"));
    if (cudaHostAlloc(reinterpret_cast<void**>(&e->h_status), 4 * sizeof(int), cudaHostAllocMapped) != cudaSuccess ||
        cudaHostGetDevicePointer(reinterpret_cast<void**>(&e->d_status), e->h_status, 0) != cudaSuccess ||
        cudaMalloc(reinterpret_cast<void**>(&e->d_max_cycles), sizeof(unsigned long long)) != cudaSuccess ||
        cudaMemset(e->d_max_cycles, 0, sizeof(unsigned long long)) != cudaSuccess)
        return (fail(CPB200_E_CUDA, "status block allocation failed: %s", cudaGetErrorString(cudaGetLastError())));
    memset(e->h_status, 0, 4 * sizeof(int));
#define UM_OPT_IN(PROF, SPLIT, MODE) (cudaFuncSetAttribute(cpb::fused_mlp_umma_kernel<PROF, SPLIT, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, cpb::UM_SMEM_BYTES) != cudaSuccess)
    if (UM_OPT_IN(false, false, 0) || UM_OPT_IN(false, false, 1) || UM_OPT_IN(false, false, 2) || UM_OPT_IN(false, true, 0) ||
        UM_OPT_IN(false, true, 1) || UM_OPT_IN(false, true, 2) || UM_OPT_IN(true, false, 0) || UM_OPT_IN(true, false, 1) ||
        UM_OPT_IN(true, false, 2) || UM_OPT_IN(true, true, 0) || UM_OPT_IN(true, true, 1) || UM_OPT_IN(true, true, 2))
        return (fail(CPB200_E_CUDA, "cannot opt in to %d bytes of shared memory: %s", cpb::UM_SMEM_BYTES,
                         cudaGetErrorString(cudaGetLastError())));
    return 0;
}
}  // namespace

extern "C" {

int cpb200_abi_version(void) { return CPB200_ABI_VERSION; }

const char* cpb200_last_error(void) { return g_err.c_str(); }

int cpb200_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    int ok = 0;
    for (int i = 0; i < n; ++i) {
        int major = 0;
        if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, i) == cudaSuccess && major == 10) ++ok;
    }
    return ok;
}

static int check_desc(const cpb200_model_desc* d) {
    if (!d) return fail(CPB200_E_INVALID, "null model description");
    if (d->n_dense < 1 || d->n_parameters < 1 || d->n_modes < 1 || !d->dims || !d->W || !d->b ||
        !d->parameters_mean || !d->parameters_std || !d->features_mean || !d->features_std)
        return fail(CPB200_E_INVALID, "incomplete model description");
    if (d->n_dense > 1 && (!d->alphas || !d->betas)) return fail(CPB200_E_INVALID, "missing activation parameters");
    if (d->dims[0] != d->n_parameters) return fail(CPB200_E_INVALID, "architecture[0] != n_parameters");
    if (d->n_pcas > 0) {
        if (!d->pca_mean || !d->pca_std || !d->pca_transform_matrix) return fail(CPB200_E_INVALID, "missing PCA arrays");
        if (d->dims[d->n_dense] != d->n_pcas) return fail(CPB200_E_INVALID, "architecture[-1] != n_pcas");
    } else if (d->dims[d->n_dense] != d->n_modes) {
        return fail(CPB200_E_INVALID, "architecture[-1] != n_modes");
    }
    for (int i = 0; i <= d->n_dense; ++i)
        if (d->dims[i] < 1) return fail(CPB200_E_INVALID, "non-positive layer width");
    return 0;
}

static int create_impl(const cpb200_model_desc* const* descs, int n_models, int device, int precision, cpb200_engine** out) {
    if (!descs || !out) return fail(CPB200_E_INVALID, "null argument");
    *out = nullptr;
    if (precision != CPB200_PREC_FP32_SIMT && precision != CPB200_PREC_F16X3)
        return fail(CPB200_E_INVALID, "unknown precision %d", precision);
    if (n_models < 1 || n_models > cpb::UM_MAX_MODELS) return fail(CPB200_E_INVALID, "1 .. %d emulators per engine", cpb::UM_MAX_MODELS);
    if (n_models > 1 && precision != CPB200_PREC_F16X3) return fail(CPB200_E_UNSUPPORTED, "multi-emulator engines exist for CPB200_PREC_F16X3 only");
    int n_gemms = 0;
    for (int m = 0; m < n_models; ++m) {
        if (int rc = check_desc(descs[m])) return rc;
        n_gemms += descs[m]->n_dense + (descs[m]->n_pcas > 0 ? 1 : 0);
    }
    if (n_gemms > cpb::UM_MAX_GEMMS) return fail(CPB200_E_UNSUPPORTED, "more than %d dense contractions", cpb::UM_MAX_GEMMS);

    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        return fail(CPB200_E_NO_DEVICE, "no CUDA device visible; this engine has no CPU fallback");
    }
    if (device < 0 || device >= ndev) return fail(CPB200_E_INVALID, "device %d out of range (%d visible)", device, ndev);
    cudaDeviceProp prop;
    CU(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10)
        return fail(CPB200_E_NO_DEVICE, "device %d is sm_%d%d; this engine is built for sm_100a only", device, prop.major, prop.minor);

    DeviceGuard guard(device);
    cpb200_engine* e = new cpb200_engine();
    e->device = device;
    e->precision = precision;
    e->n_models = n_models;
    e->P = descs[0]->n_parameters;
    e->n_modes = descs[0]->n_modes;
    e->n_pcas = descs[0]->n_pcas;
    e->num_sms = prop.multiProcessorCount;
    auto bail = [&](int code) { cpb200_destroy(e); return code; };
    int rc = 0;
    for (int m = 0; m < n_models; ++m) {
        e->mP[m] = descs[m]->n_parameters;
        e->mModes[m] = descs[m]->n_modes;
        if ((rc = build_gemms(descs[m], e->gemms, m))) return bail(rc);
        if ((rc = upload(&e->md_pmean[m], descs[m]->parameters_mean, e->mP[m]))) return bail(rc);
        if ((rc = upload(&e->md_pstd[m], descs[m]->parameters_std, e->mP[m]))) return bail(rc);
    }
    e->d_pmean = e->md_pmean[0];
    e->d_pstd = e->md_pstd[0];

    if (precision == CPB200_PREC_FP32_SIMT) {
        int width = 1;
        for (auto& g : e->gemms) {
            if ((rc = upload(&g.dW, g.W.data(), g.W.size()))) return bail(rc);
            if (round_up(g.K, 4) <= cpb::SMALL_KP_MAX) {      // transposed copy for dense_small_kernel
                g.Kp = round_up(g.K, 4);
                std::vector<float> wt(static_cast<size_t>(g.N) * g.Kp, 0.f);
                for (int k = 0; k < g.K; ++k)
                    for (int n = 0; n < g.N; ++n) wt[static_cast<size_t>(n) * g.Kp + k] = g.W[static_cast<size_t>(k) * g.N + n];
                if ((rc = upload(&g.dWt, wt.data(), wt.size()))) return bail(rc);
            }
            if ((rc = upload(&g.dp0, g.p0.data(), g.p0.size()))) return bail(rc);
            if ((rc = upload(&g.dp1, g.p1.data(), g.p1.size()))) return bail(rc);
            if ((rc = upload(&g.dp2, g.p2.data(), g.p2.size()))) return bail(rc);
            if (&g != &e->gemms.back()) width = std::max(width, g.N);
        }
        e->simt_width = width;
        e->simt_rows = 32768;
        for (int i = 0; i < 2; ++i)
            if (cudaMalloc(reinterpret_cast<void**>(&e->simt_buf[i]), sizeof(float) * e->simt_rows * width) != cudaSuccess)
                return bail(fail(CPB200_E_CUDA, "scratch allocation failed: %s", cudaGetErrorString(cudaGetLastError())));
    } else {
        if ((rc = setup_umma(e))) return bail(rc);
    }
    for (auto& g : e->gemms) { g.W.clear(); g.W.shrink_to_fit(); }
    if (cudaEventCreate(&e->ev0) != cudaSuccess || cudaEventCreate(&e->ev1) != cudaSuccess)
        return bail(fail(CPB200_E_CUDA, "event creation failed"));
    if (cudaDeviceSynchronize() != cudaSuccess)
        return bail(fail(CPB200_E_CUDA, "upload failed: %s", cudaGetErrorString(cudaGetLastError())));
    *out = e;
    return 0;
}

int cpb200_create(const cpb200_model_desc* d, int device, int precision, cpb200_engine** out) {
    if (!d || !out) return fail(CPB200_E_INVALID, "null argument");
    return create_impl(&d, 1, device, precision, out);
}

int cpb200_create_multi(const cpb200_model_desc* const* descs, int32_t n_models, int device, cpb200_engine** out) {
    return create_impl(descs, n_models, device, CPB200_PREC_F16X3, out);
}

int cpb200_forward_multi_device(cpb200_engine* e, const float* const* params_dev, int64_t n_rows, float* const* out_dev,
                                const int64_t* out_pitch, const int32_t* flags, void* stream) {
    if (!e) return fail(CPB200_E_INVALID, "null engine");
    if (n_rows < 0) return fail(CPB200_E_INVALID, "negative row count");
    if (n_rows == 0) return 0;
    if (!params_dev || !out_dev || !out_pitch || !flags) return fail(CPB200_E_INVALID, "null argument");
    int fl[cpb::UM_MAX_MODELS];
    for (int m = 0; m < e->n_models; ++m) {
        if (!params_dev[m] || !out_dev[m]) return fail(CPB200_E_INVALID, "null buffer for emulator %d", m);
        if (out_pitch[m] < e->mModes[m]) return fail(CPB200_E_INVALID, "out_pitch %lld < n_modes %d (emulator %d)", static_cast<long long>(out_pitch[m]), e->mModes[m], m);
        fl[m] = flags[m];
        if (flags[m] & CPB200_FLAG_BINNED) return fail(CPB200_E_INVALID, "CPB200_FLAG_BINNED is a single-emulator option");
        if (flags[m] & CPB200_FLAG_ACCUMULATE) {
            // the sum is read back by the very lane that wrote it (same chunking): emulator m adds to what emulator m-1 left in
            // the same buffer, a few microseconds earlier and still in L2
            if (m == 0 || out_dev[m] != out_dev[m - 1] || out_pitch[m] != out_pitch[m - 1] || e->mModes[m] != e->mModes[m - 1])
                return fail(CPB200_E_INVALID, "CPB200_FLAG_ACCUMULATE on emulator %d needs the previous emulator's output buffer, pitch and n_modes", m);
            if (flags[m - 1] & CPB200_FLAG_TEN_TO) return fail(CPB200_E_INVALID, "the emulator before an accumulating one must leave log10 values (no CPB200_FLAG_TEN_TO)");
        }
    }
    DeviceGuard guard(e->device);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (e->timing) CU(cudaEventRecord(e->ev0, st));
    if (int rc = order_before(e, st)) return rc;
    if (int rc = launch_umma(e, params_dev, n_rows, out_dev, out_pitch, fl, st)) return rc;
    if (int rc = order_after(e, st)) return rc;
    if (e->timing) { CU(cudaEventRecord(e->ev1, st)); e->ev_valid = true; }
    return 0;
}

int cpb200_destroy(cpb200_engine* e) {
    if (!e) return 0;
    DeviceGuard guard(e->device);
    cudaDeviceSynchronize();
    for (auto& g : e->gemms) {
        cudaFree(g.dW); cudaFree(g.dWt); cudaFree(g.dp0); cudaFree(g.dp1); cudaFree(g.dp2);
        cudaFree(g.dwpack); cudaFree(g.depi); cudaFree(g.depi_b);
    }
    for (int m = 0; m < cpb::UM_MAX_MODELS; ++m) { cudaFree(e->md_pmean[m]); cudaFree(e->md_pstd[m]); }
    cudaFree(e->simt_buf[0]); cudaFree(e->simt_buf[1]);
    cudaFree(e->um_scratch); cudaFree(e->d_max_cycles); cudaFree(e->d_counters);
    if (e->h_status) cudaFreeHost(e->h_status);
    if (e->ev_order) cudaEventDestroy(e->ev_order);
    for (int i = 0; i < 2; ++i) {
        if (e->hstream[i]) cudaStreamDestroy(e->hstream[i]);
        if (e->h_in[i]) cudaFreeHost(e->h_in[i]);
        if (e->h_out[i]) cudaFreeHost(e->h_out[i]);
        cudaFree(e->d_in[i]); cudaFree(e->d_out[i]);
    }
    for (cudaEvent_t ev : e->ev_piece) if (ev) cudaEventDestroy(ev);
    if (e->ev0) cudaEventDestroy(e->ev0);
    if (e->ev1) cudaEventDestroy(e->ev1);
    cudaGetLastError();
    delete e;
    return 0;
}

int cpb200_forward_device(cpb200_engine* e, const float* params_dev, int64_t n_rows, float* out_dev,
                          int64_t out_pitch, int flags, void* stream) {
    if (!e) return fail(CPB200_E_INVALID, "null engine");
    if (n_rows < 0) return fail(CPB200_E_INVALID, "negative row count");
    if (n_rows == 0) return 0;
    if (!params_dev || !out_dev) return fail(CPB200_E_INVALID, "null buffer");
    if (e->n_models != 1) return fail(CPB200_E_INVALID, "a multi-emulator engine is driven through cpb200_forward_multi_device");
    if (flags & CPB200_FLAG_BINNED) {
        const Gemm& last = e->gemms.back();
        if (e->precision != CPB200_PREC_F16X3 || !last.depi_b) return fail(CPB200_E_INVALID, "CPB200_FLAG_BINNED needs cpb200_engine_set_binning first");
        if (!(flags & CPB200_FLAG_TEN_TO) || (flags & CPB200_FLAG_ACCUMULATE)) return fail(CPB200_E_INVALID, "CPB200_FLAG_BINNED bins 10**y: it goes with CPB200_FLAG_TEN_TO, without ACCUMULATE");
        const int64_t need = cpb200_engine_binned_width(e);
        if (out_pitch < need || (out_pitch & 3) || (reinterpret_cast<uintptr_t>(out_dev) & 15))
            return fail(CPB200_E_INVALID, "binned output rows are %lld floats; the buffer must be 16-byte aligned with a pitch that is a multiple of 4", static_cast<long long>(need));
    } else
    if (out_pitch < e->n_modes) return fail(CPB200_E_INVALID, "out_pitch %lld < n_modes %d", static_cast<long long>(out_pitch), e->n_modes);
    DeviceGuard guard(e->device);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (e->timing) CU(cudaEventRecord(e->ev0, st));
    int rc = forward_async(e, params_dev, n_rows, out_dev, out_pitch, flags, st);
    if (rc) return rc;
    if (e->timing) { CU(cudaEventRecord(e->ev1, st)); e->ev_valid = true; }
    return 0;
}

// (the "ready" markers are set only once every allocation has succeeded; a partial set-up is retried - what was
// allocated is reused - and released by cpb200_destroy)
static int ensure_host_pipeline(cpb200_engine* e) {
    if (e->pipeline_ready) return 0;
    e->chunk_rows = std::max<int64_t>(256, HOST_CHUNK_BYTES / (static_cast<int64_t>(e->n_modes) * 4) / 256 * 256);
    e->dpitch = (e->n_modes + 7) / 8 * 8;             // 32-byte rows: every 8-byte pair store stays inside one sector
    for (int i = 0; i < 2; ++i) {
        if (!e->hstream[i]) CU(cudaStreamCreateWithFlags(&e->hstream[i], cudaStreamNonBlocking));
        if (!e->d_in[i]) CU(cudaMalloc(reinterpret_cast<void**>(&e->d_in[i]), sizeof(float) * e->chunk_rows * e->P));
        if (!e->d_out[i]) CU(cudaMalloc(reinterpret_cast<void**>(&e->d_out[i]), sizeof(float) * e->chunk_rows * e->dpitch));
    }
    e->pipeline_ready = true;
    return 0;
}

static int ensure_bounce(cpb200_engine* e) {
    if (e->bounce_ready) return 0;
    for (int i = 0; i < 2; ++i) {
        if (!e->h_in[i]) CU(cudaMallocHost(reinterpret_cast<void**>(&e->h_in[i]), sizeof(float) * e->chunk_rows * e->P));
        if (!e->h_out[i]) CU(cudaMallocHost(reinterpret_cast<void**>(&e->h_out[i]), sizeof(float) * e->chunk_rows * e->n_modes));
    }
    e->bounce_ready = true;
    return 0;
}

static bool is_pinned(const void* p) {
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return a.type == cudaMemoryTypeHost;
}

}  // extern "C"

namespace {
// Host threads used to copy results out of the pinned bounce buffers (CPB200_HOST_THREADS overrides; at most 32).
int host_threads() {
    static const int n = [] {
        const char* s = getenv("CPB200_HOST_THREADS");
        int v = s ? atoi(s) : static_cast<int>(std::thread::hardware_concurrency());
        if (!s) {   // no more than the CPUs this process may run on (taskset, cpusets, a rank pinned next to its GPU)
            cpu_set_t set;
            CPU_ZERO(&set);
            if (sched_getaffinity(0, sizeof set, &set) == 0 && CPU_COUNT(&set) > 0) v = std::min(v, CPU_COUNT(&set));
        }
        return std::max(1, std::min(v, 32));
    }();
    return n;
}

// A small persistent pool of host threads for the copy-out below.  Spawning threads per call is slow on this pool's guests
// (~100 us apiece), and so is waking a sleeping one (tools/latency_breakdown.py, CPB200_HOST_THREADS = 1 / 4 / 16 on a 10 MB
// result: 1.12 / 1.02 / 1.53 ms).  Hence SESSIONS: a host call that will copy out a few megabytes or more opens one before it
// launches - the workers wake while the kernel and the transfer run and then poll a generation counter instead of sleeping,
// so each piece of the result finds them ready - and closes it when it returns (RAII), which sends them back to sleep.  One
// session at a time: a caller that finds the pool taken (another device's host thread) copies on its own.
class HostPool {
public:
    static HostPool* get() {
        static std::mutex guard;
        static HostPool* pool = nullptr;
        static pid_t owner = 0;
        std::lock_guard<std::mutex> lk(guard);
        if (!pool || owner != getpid()) {          // (a forked child has no workers: it gets a pool of its own; the old object leaks)
            pool = new HostPool(std::max(0, host_threads() - 1));
            owner = getpid();
        }
        return pool;
    }
    class Session {
    public:
        // `helpers`: how many workers to wake (0 = no session) - only as many as the result can feed, the others stay asleep
        explicit Session(int helpers) {
            if (helpers <= 0) return;
            HostPool* p = HostPool::get();
            if (p->workers_ == 0 || !p->use_.try_lock()) return;
            pool_ = p;
            p->awake_.store(std::min(helpers, p->workers_), std::memory_order_release);
            p->hot_.store(true, std::memory_order_release);
            { std::lock_guard<std::mutex> lk(p->m_); }
            p->cv_work_.notify_all();
        }
        ~Session() {
            if (!pool_) return;
            pool_->hot_.store(false, std::memory_order_release);
            pool_->use_.unlock();
        }
        Session(const Session&) = delete;
        Session& operator=(const Session&) = delete;
        bool active() const { return pool_ != nullptr; }
        int threads() const { return pool_ ? pool_->awake_.load(std::memory_order_relaxed) + 1 : 1; }
        // f(i) for i in [0, n) on the workers and the calling thread; returns when all are done
        void run(int n, const std::function<void(int)>& f) {
            HostPool* p = pool_;
            {
                std::lock_guard<std::mutex> lk(p->m_);
                p->job_ = &f; p->n_ = n; p->next_ = 0;
                p->pending_.store(n, std::memory_order_relaxed);
            }
            p->gen_.fetch_add(1, std::memory_order_release);
            p->work();
            while (p->pending_.load(std::memory_order_acquire) != 0) cpu_relax();
            std::lock_guard<std::mutex> lk(p->m_);
            p->job_ = nullptr;
        }
    private:
        HostPool* pool_ = nullptr;
    };
private:
    static void cpu_relax() {
#if defined(__x86_64__) || defined(__i386__)
        __builtin_ia32_pause();
#else
        std::this_thread::yield();
#endif
    }
    explicit HostPool(int n_workers) : workers_(n_workers) {
        for (int i = 0; i < n_workers; ++i) std::thread([this, i] { loop(i); }).detach();      // they live as long as the process
    }
    // claim and run tasks of the current job until none is left (task claims go through the mutex: tasks are coarse)
    void work() {
        for (;;) {
            int i;
            const std::function<void(int)>* f;
            {
                std::lock_guard<std::mutex> lk(m_);
                if (!job_ || next_ >= n_) return;
                i = next_++; f = job_;
            }
            (*f)(i);
            pending_.fetch_sub(1, std::memory_order_release);
        }
    }
    void loop(int index) {
        uint64_t seen = gen_.load(std::memory_order_acquire);
        unsigned idle = 0;
        auto wanted = [&] { return hot_.load(std::memory_order_acquire) && index < awake_.load(std::memory_order_acquire); };
        for (;;) {
            if (!wanted()) {
                std::unique_lock<std::mutex> lk(m_);
                cv_work_.wait(lk, wanted);
            }
            const uint64_t g = gen_.load(std::memory_order_acquire);
            if (g != seen) { seen = g; work(); idle = 0; }
            else if (++idle < 4096) cpu_relax();
            else std::this_thread::yield();          // an oversubscribed host (CPU quota, many ranks): give the caller the core
        }
    }
    const int workers_;
    std::mutex m_, use_;
    std::condition_variable cv_work_;
    std::atomic<bool> hot_{false};
    std::atomic<int> awake_{0};                           // workers 0 .. awake_ - 1 take part in the current session
    std::atomic<uint64_t> gen_{0};
    std::atomic<int> pending_{0};
    const std::function<void(int)>* job_ = nullptr;      // guarded by m_
    int n_ = 0, next_ = 0;                                // guarded by m_
};

// rows of float32 in a pinned bounce buffer -> the caller's (pageable) rows of float32 or float64.  Fresh NumPy arrays are
// untouched memory: the page faults of the first write dominate a single-threaded copy (3-4 GB/s), so the rows are split
// over the session's threads - at least 256 KB of input per thread; without a session the calling thread copies alone.
template <typename T>
void copy_rows_out(T* dst, int64_t dst_pitch, const float* src, int64_t src_pitch, int64_t rows, int cols, HostPool::Session& sess) {
    auto work = [=](int64_t r0, int64_t r1) {
        for (int64_t r = r0; r < r1; ++r) {
            const float* s = src + r * src_pitch;
            T* d = dst + r * dst_pitch;
            if (std::is_same<T, float>::value) memcpy(d, s, sizeof(float) * cols);
            else for (int j = 0; j < cols; ++j) d[j] = static_cast<T>(s[j]);
        }
    };
    if (sess.active()) {
        const int nt = static_cast<int>(std::min<int64_t>(sess.threads(), rows * cols / (1 << 16)));
        if (nt <= 1) { work(0, rows); return; }
        sess.run(nt, [&](int t) { work(rows * t / nt, rows * (t + 1) / nt); });
        return;
    }
    // no session (the pool serves another device's host thread, or the result is small): threads of its own from 8 MB on,
    // where their start-up cost is a small part of the copy
    const int nt = rows * static_cast<int64_t>(cols) >= (2 << 20) ? static_cast<int>(std::min<int64_t>(host_threads(), rows * cols / (1 << 18))) : 1;
    if (nt <= 1) { work(0, rows); return; }
    std::vector<std::thread> pool;
    pool.reserve(nt);
    for (int t = 0; t < nt; ++t) pool.emplace_back(work, rows * t / nt, rows * (t + 1) / nt);
    for (std::thread& th : pool) th.join();
}

template <typename T>
int forward_host_impl(cpb200_engine* e, const float* params_host, int64_t n_rows, T* out_host, int64_t out_pitch, int flags) {
    if (!e) return fail(CPB200_E_INVALID, "null engine");
    if (n_rows < 0) return fail(CPB200_E_INVALID, "negative row count");
    if (n_rows == 0) return 0;
    if (!params_host || !out_host) return fail(CPB200_E_INVALID, "null buffer");
    if (e->n_models != 1) return fail(CPB200_E_INVALID, "a multi-emulator engine is driven through cpb200_forward_multi_device");
    if (out_pitch < e->n_modes) return fail(CPB200_E_INVALID, "out_pitch %lld < n_modes %d", static_cast<long long>(out_pitch), e->n_modes);
    if (flags & CPB200_FLAG_ACCUMULATE)
        return fail(CPB200_E_INVALID, "CPB200_FLAG_ACCUMULATE needs the first emulator's result on the device: device path only");
    DeviceGuard guard(e->device);
    if (int rc = ensure_host_pipeline(e)) return rc;
    // float32 results in pinned memory are written by the DMA engine directly; anything else is copied (and widened)
    // out of the pinned bounce buffers by the host threads while the next chunk is in flight
    const bool pin_in = is_pinned(params_host), pin_out = std::is_same<T, float>::value && is_pinned(out_host);
    if (!(pin_in && pin_out)) if (int rc = ensure_bounce(e)) return rc;
    // chunk size: the engine's (128 MB of output) for direct DMA; with a host copy behind it, about an eighth of the
    // batch (>= 2048 rows), so that the copy-out of one chunk overlaps the transfer of the next
    int64_t R = e->chunk_rows;
    if (!pin_out) R = std::min<int64_t>(R, std::max<int64_t>(2048, (n_rows + 7) / 8 + 255) / 256 * 256);
    const int64_t n_chunks = (n_rows + R - 1) / R;
    const size_t row_bytes = sizeof(float) * e->n_modes;
    // helpers for the copy-out: woken now, ready when the first piece lands (4 MB of result or more; see HostPool)
    // (one helper per 2 MB of result: a pool that wakes wider than the copy can use only competes with the caller's other threads)
    const int64_t result_bytes = n_rows * static_cast<int64_t>(sizeof(T)) * e->n_modes;
    HostPool::Session session(!pin_out && result_bytes >= (4ll << 20) ? static_cast<int>(std::min<int64_t>(64, result_bytes >> 21)) : 0);
    // Tight host rows (the NumPy layout, and the bounce buffers): the kernel then writes tight rows on the device too and the
    // chunk leaves as ONE contiguous copy - 56.9 GB/s on this pool's boxes against 52.4 GB/s for the pitched 2-D copy of
    // 10028-byte rows (tools/microbench/d2h_copy_bw.cu).  Unaligned rows cost the kernel its vector stores (cmb_TT: 2x the
    // kernel time), but a chunk's kernel (well under a millisecond) hides behind the previous chunk's transfer (2.4 ms).
    // Only for transfers that PCIe dominates (>= 256 MB): a mid-size call is a handful of short launches whose latency counts.
    const bool tight = (!pin_out || out_pitch == e->n_modes) && n_rows * static_cast<int64_t>(row_bytes) >= (256ll << 20);
    const int64_t dev_pitch = tight ? e->n_modes : e->dpitch;
    // two slots, each with its own stream: H2D -> kernel -> D2H; slot reuse waits for its stream.  The kernels of the two
    // slots (and any earlier device-path launch of this engine) are ordered by forward_async: they share the scratch.
    // One chunk's submission; a failure leaves through `bail`, which drains both streams first (no copy stays in flight
    // into the caller's buffers).
    auto submit = [&](int64_t c, int slot) -> int {
        const int64_t r0 = c * R, m = std::min<int64_t>(R, n_rows - r0);
        const float* src = params_host + r0 * e->P;
        if (!pin_in) { memcpy(e->h_in[slot], src, sizeof(float) * m * e->P); src = e->h_in[slot]; }
        CU(cudaMemcpyAsync(e->d_in[slot], src, sizeof(float) * m * e->P, cudaMemcpyHostToDevice, e->hstream[slot]));
        if (int rc = forward_async(e, e->d_in[slot], m, e->d_out[slot], dev_pitch, flags, e->hstream[slot])) return rc;
        if (tight) {
            float* dst = pin_out ? reinterpret_cast<float*>(out_host) + r0 * out_pitch : e->h_out[slot];
            CU(cudaMemcpyAsync(dst, e->d_out[slot], row_bytes * m, cudaMemcpyDeviceToHost, e->hstream[slot]));
        } else if (pin_out) {
            CU(cudaMemcpy2DAsync(reinterpret_cast<float*>(out_host) + r0 * out_pitch, sizeof(float) * out_pitch, e->d_out[slot],
                                 sizeof(float) * e->dpitch, row_bytes, m, cudaMemcpyDeviceToHost, e->hstream[slot]));
        } else {
            CU(cudaMemcpy2DAsync(e->h_out[slot], row_bytes, e->d_out[slot], sizeof(float) * e->dpitch, row_bytes, m,
                                 cudaMemcpyDeviceToHost, e->hstream[slot]));
        }
        return 0;
    };
    auto bail = [&](int rc) {
        const std::string msg = g_err;                      // the drain below must not replace the first error's message
        cudaStreamSynchronize(e->hstream[0]);
        cudaStreamSynchronize(e->hstream[1]);
        cudaGetLastError();
        g_err = msg;
        return rc;
    };
    // Mid-size pageable result (one chunk, a megabyte or more): ONE launch, then the transfer in up to 16 pieces, each copied
    // out of the bounce buffer by this thread while the next ones are still on the wire - the copy-out (a single thread
    // moves ~17 GB/s here; from 4 MB on the session's helpers share it) hides the transfer instead of following it
    // (1024 cmb_TT rows: 1.1-1.5 ms before; tools/latency_breakdown.py).
    if (!pin_out && n_chunks == 1 && n_rows * static_cast<int64_t>(row_bytes) >= (1ll << 20)) {
        const int n_pieces = static_cast<int>(std::min<int64_t>(16, std::max<int64_t>(1, n_rows * static_cast<int64_t>(row_bytes) / (1ll << 20))));
        for (int i = 0; i < n_pieces; ++i) if (!e->ev_piece[i]) CU(cudaEventCreateWithFlags(&e->ev_piece[i], cudaEventDisableTiming));
        const float* src = params_host;
        if (!pin_in) { memcpy(e->h_in[0], src, sizeof(float) * n_rows * e->P); src = e->h_in[0]; }
        CU(cudaMemcpyAsync(e->d_in[0], src, sizeof(float) * n_rows * e->P, cudaMemcpyHostToDevice, e->hstream[0]));
        if (int rc = forward_async(e, e->d_in[0], n_rows, e->d_out[0], dev_pitch, flags, e->hstream[0])) return bail(rc);
        auto piece_lo = [&](int i) { return n_rows * i / n_pieces; };
        for (int i = 0; i < n_pieces; ++i) {
            const int64_t lo = piece_lo(i), m = piece_lo(i + 1) - lo;
            const cudaError_t err = cudaMemcpy2DAsync(e->h_out[0] + lo * e->n_modes, row_bytes, e->d_out[0] + lo * dev_pitch, sizeof(float) * dev_pitch,
                                                      row_bytes, m, cudaMemcpyDeviceToHost, e->hstream[0]);
            if (err != cudaSuccess || cudaEventRecord(e->ev_piece[i], e->hstream[0]) != cudaSuccess)
                return bail(fail(CPB200_E_CUDA, "device->host piece %d failed: %s", i, cudaGetErrorString(cudaGetLastError())));
        }
        for (int i = 0; i < n_pieces; ++i) {
            const cudaError_t err = cudaEventSynchronize(e->ev_piece[i]);
            if (err != cudaSuccess) return bail(fail_sync(e, err, "cudaEventSynchronize"));
            const int64_t lo = piece_lo(i), m = piece_lo(i + 1) - lo;
            copy_rows_out(out_host + lo * out_pitch, out_pitch, e->h_out[0] + lo * e->n_modes, e->n_modes, m, e->n_modes, session);
        }
        return check_range(e);
    }
    for (int64_t c = 0; c < n_chunks + 2; ++c) {
        const int slot = static_cast<int>(c & 1);
        if (c >= 2) {   // retire chunk c-2 of this slot
            const cudaError_t err = cudaStreamSynchronize(e->hstream[slot]);
            if (err != cudaSuccess) return bail(fail_sync(e, err, "cudaStreamSynchronize"));
            const int64_t r0 = (c - 2) * R, m = std::min<int64_t>(R, n_rows - r0);
            if (!pin_out) copy_rows_out(out_host + r0 * out_pitch, out_pitch, e->h_out[slot], e->n_modes, m, e->n_modes, session);
        }
        if (c < n_chunks) if (int rc = submit(c, slot)) return bail(rc);
    }
    return check_range(e);
}
}  // namespace

extern "C" {

int cpb200_forward_host(cpb200_engine* e, const float* params_host, int64_t n_rows, float* out_host,
                        int64_t out_pitch, int flags) {
    return forward_host_impl<float>(e, params_host, n_rows, out_host, out_pitch, flags);
}

int cpb200_forward_host_f64(cpb200_engine* e, const float* params_host, int64_t n_rows, double* out_host,
                            int64_t out_pitch, int flags) {
    return forward_host_impl<double>(e, params_host, n_rows, out_host, out_pitch, flags);
}

}  // extern "C"

namespace {
// CPUs next to a CUDA device (sysfs local_cpulist of its PCI function); empty when the kernel does not say
std::vector<int> device_local_cpus(int device) {
    std::vector<int> cpus;
    char bus[32] = {0};
    if (cudaDeviceGetPCIBusId(bus, sizeof bus, device) != cudaSuccess) { cudaGetLastError(); return cpus; }
    for (char* c = bus; *c; ++c) *c = static_cast<char>(tolower(*c));
    const std::string path = std::string("/sys/bus/pci/devices/") + bus + "/local_cpulist";
    FILE* f = fopen(path.c_str(), "r");
    if (!f) return cpus;
    char line[4096] = {0};
    if (fgets(line, sizeof line, f)) {
        for (char* tok = strtok(line, ",\n"); tok; tok = strtok(nullptr, ",\n")) {
            int a = 0, b = 0;
            const int got = sscanf(tok, "%d-%d", &a, &b);
            if (got == 1) b = a;
            if (got >= 1) for (int c = a; c <= b && c < CPU_SETSIZE; ++c) cpus.push_back(c);
        }
    }
    fclose(f);
    return cpus;
}

// Bind the calling thread to the CPUs next to `device` (those it is allowed to use): the pinned bounce buffers it
// allocates and the result pages it first touches then sit on the GPU's NUMA node.
void bind_thread_to_device(int device) {
    const std::vector<int> cpus = device_local_cpus(device);
    if (cpus.empty()) return;
    cpu_set_t allowed, want;
    CPU_ZERO(&allowed); CPU_ZERO(&want);
    if (pthread_getaffinity_np(pthread_self(), sizeof allowed, &allowed) != 0) return;
    int n = 0;
    for (int c : cpus) if (CPU_ISSET(c, &allowed)) { CPU_SET(c, &want); ++n; }
    if (n > 0) pthread_setaffinity_np(pthread_self(), sizeof want, &want);
}

template <typename T>
int forward_host_multi_impl(cpb200_engine* const* engines, int n_engines, const float* params_host, int64_t n_rows,
                            T* out_host, int64_t out_pitch, int flags) {
    if (!engines || n_engines < 1) return fail(CPB200_E_INVALID, "no engines");
    for (int i = 0; i < n_engines; ++i) {
        if (!engines[i]) return fail(CPB200_E_INVALID, "null engine");
        if (engines[i]->P != engines[0]->P || engines[i]->n_modes != engines[0]->n_modes)
            return fail(CPB200_E_INVALID, "engines describe different models");
        for (int j = 0; j < i; ++j) if (engines[j] == engines[i]) return fail(CPB200_E_INVALID, "engine listed twice");
    }
    if (n_rows < 0) return fail(CPB200_E_INVALID, "negative row count");
    if (n_engines == 1 || n_rows < 2 * n_engines) return forward_host_impl<T>(engines[0], params_host, n_rows, out_host, out_pitch, flags);
    // contiguous row shards, sizes differing by at most one row (cosmopower_b200/sharding.py shard_bounds); one host
    // thread per device, each running the ordinary pipelined host path of its engine into its slice of the result
    std::vector<int> rc(n_engines, 0);
    std::vector<std::string> msg(n_engines);
    std::vector<std::thread> pool;
    const int64_t base = n_rows / n_engines, extra = n_rows % n_engines;
    for (int i = 0; i < n_engines; ++i) {
        const int64_t lo = i * base + std::min<int64_t>(i, extra), m = base + (i < extra ? 1 : 0);
        pool.emplace_back([=, &rc, &msg] {
            bind_thread_to_device(engines[i]->device);
            rc[i] = forward_host_impl<T>(engines[i], params_host + lo * engines[i]->P, m, out_host + lo * out_pitch, out_pitch, flags);
            if (rc[i]) msg[i] = g_err;
        });
    }
    for (std::thread& th : pool) th.join();
    for (int i = 0; i < n_engines; ++i)
        if (rc[i] && rc[i] != CPB200_E_RANGE) return fail(rc[i], "device %d: %s", engines[i]->device, msg[i].c_str());
    for (int i = 0; i < n_engines; ++i)
        if (rc[i]) return fail(rc[i], "device %d: %s", engines[i]->device, msg[i].c_str());
    return 0;
}
}  // namespace

extern "C" {

int cpb200_forward_host_multi(cpb200_engine* const* engines, int32_t n_engines, const float* params_host, int64_t n_rows,
                              float* out_host, int64_t out_pitch, int flags) {
    return forward_host_multi_impl<float>(engines, n_engines, params_host, n_rows, out_host, out_pitch, flags);
}

int cpb200_forward_host_multi_f64(cpb200_engine* const* engines, int32_t n_engines, const float* params_host, int64_t n_rows,
                                  double* out_host, int64_t out_pitch, int flags) {
    return forward_host_multi_impl<double>(engines, n_engines, params_host, n_rows, out_host, out_pitch, flags);
}

int cpb200_host_alloc(void** ptr, uint64_t bytes) {
    if (!ptr) return fail(CPB200_E_INVALID, "null argument");
    CU(cudaMallocHost(ptr, bytes ? bytes : 1));
    return 0;
}

int cpb200_host_free(void* ptr) {
    if (ptr) CU(cudaFreeHost(ptr));
    return 0;
}

int cpb200_engine_info(const cpb200_engine* e, int32_t* n_parameters, int32_t* n_modes, int32_t* precision,
                       int32_t* device) {
    if (!e) return fail(CPB200_E_INVALID, "null engine");
    if (n_parameters) *n_parameters = e->P;
    if (n_modes) *n_modes = e->n_modes;
    if (precision) *precision = e->precision;
    if (device) *device = e->device;
    return 0;
}

int cpb200_engine_debug_counters(cpb200_engine* e, int enable, int64_t* out, int32_t capacity) {
    if (!e) return fail(CPB200_E_INVALID, "null engine");
    DeviceGuard guard(e->device);
    const size_t n = static_cast<size_t>(e->num_sms + cpb::UM_MAX_JOBS) * cpb::UM_CNT_SLOTS      // per-CTA rows, then per-job rows,
                     + static_cast<size_t>(cpb::UM_MAX_JOBS) * cpb::UM_TRACE_STAGES * 2;                // then the stage trace
    if (out && e->d_counters) {
        CU(cudaDeviceSynchronize());
        std::vector<long long> h(n);
        CU(cudaMemcpy(h.data(), e->d_counters, n * sizeof(long long), cudaMemcpyDeviceToHost));
        for (size_t i = 0; i < n && i < static_cast<size_t>(capacity); ++i) out[i] = h[i];
    }
    if (enable && !e->d_counters) {
        CU(cudaMalloc(reinterpret_cast<void**>(&e->d_counters), n * sizeof(long long)));
        CU(cudaMemset(e->d_counters, 0, n * sizeof(long long)));
    } else if (!enable && e->d_counters) {
        CU(cudaDeviceSynchronize());
        cudaFree(e->d_counters);
        e->d_counters = nullptr;
    }
    return static_cast<int>(n);
}

int cpb200_debug_job_table(int32_t n_gemms, const int32_t* k, const int32_t* n, const int32_t* epi_kind, const int32_t* model, int32_t tri,
                           uint32_t* w0, uint32_t* w1, int32_t capacity, int32_t* info) {
    if (n_gemms < 1 || n_gemms > cpb::UM_MAX_GEMMS || !k || !n || !epi_kind) return fail(CPB200_E_INVALID, "bad job-table request");
    std::vector<Gemm> gemms(n_gemms);
    for (int i = 0; i < n_gemms; ++i) {
        gemms[i].K = k[i]; gemms[i].N = n[i]; gemms[i].epi = epi_kind[i]; gemms[i].tri = tri != 0;
        gemms[i].model = model ? model[i] : 0;
        gemms[i].layer0 = i == 0 || gemms[i].model != gemms[i - 1].model;
    }
    for (int i = 0; i < n_gemms; ++i) gemms[i].last = i + 1 == n_gemms || gemms[i + 1].model != gemms[i].model;
    int k_pad = round_up(k[0], cpb::UM_KS);
    const int sched = plan_schedule(gemms, k_pad);
    for (int i = 0; i < n_gemms; ++i) {
        if (gemms[i].layer0) k_pad = round_up(k[i], cpb::UM_KS);
        plan_chunks(gemms[i], k_pad);
        k_pad = gemms[i].n_chunks * gemms[i].nc;
    }
    std::vector<JobEntry> entries;
    int rot_mul = 0, par_stages = 0;
    if (int rc = build_jobs(gemms, sched, entries, &rot_mul, &par_stages)) return rc;
    std::vector<unsigned> a; std::vector<unsigned short> b;
    encode_jobs(entries, a, b);
    for (size_t i = 0; i < a.size() && static_cast<int>(i) < capacity; ++i) { if (w0) w0[i] = a[i]; if (w1) w1[i] = b[i]; }
    if (info) {
        info[0] = sched; info[1] = rot_mul; info[2] = par_stages;
        for (int i = 0; i < n_gemms; ++i) { info[3 + 3 * i] = gemms[i].k_stages; info[4 + 3 * i] = gemms[i].nc; info[5 + 3 * i] = gemms[i].n_chunks; }
    }
    return static_cast<int>(a.size());
}

int64_t cpb200_engine_launch_count(const cpb200_engine* e) { return e ? e->launches : 0; }

int64_t cpb200_engine_binned_width(const cpb200_engine* e) {
    if (!e) return 0;
    return round_up(2 * ((e->n_modes + 3) / 4), 4);
}

int cpb200_engine_set_binning(cpb200_engine* e, const float* weight, const int32_t* bin_of_mode) {
    if (!e || !weight || !bin_of_mode) return fail(CPB200_E_INVALID, "null argument");
    if (e->precision != CPB200_PREC_F16X3 || e->n_models != 1) return fail(CPB200_E_UNSUPPORTED, "binning is fused into single-emulator tensor-core engines");
    Gemm& g = e->gemms.back();
    if (g.epi != cpb::EPI_OUT) return fail(CPB200_E_UNSUPPORTED, "the engine has no spectrum output layer");
    // a group of four multipoles may touch two bins at most (first / second partial of the group), bins ascend with the mode
    for (int n = 0; n + 1 < g.N; ++n)
        if (bin_of_mode[n] >= 0 && bin_of_mode[n + 1] >= 0 && bin_of_mode[n + 1] < bin_of_mode[n])
            return fail(CPB200_E_INVALID, "bins must ascend with the mode index");
    std::vector<float4> epi(static_cast<size_t>(g.n_chunks) * g.nc, make_float4(0.f, 0.f, 0.f, 0.f));
    const double l2_10 = 3.321928094887362;
    for (int g0 = 0; g0 < g.N; g0 += 4) {
        int first = -1;
        for (int n = g0; n < std::min(g0 + 4, g.N); ++n) if (bin_of_mode[n] >= 0) { first = bin_of_mode[n]; break; }
        for (int n = g0; n < std::min(g0 + 4, g.N); ++n) {
            const int b = bin_of_mode[n];
            if (b >= 0 && b != first && b != first + 1) return fail(CPB200_E_UNSUPPORTED, "modes %d .. %d touch more than two bins", g0, g0 + 3);
            const int nb = n & 31, t = nb >> 3, j = nb & 7;
            float4& dst = epi[(n & ~31) + 4 * j + t];
            dst.x = static_cast<float>(static_cast<double>(g.p0[n]) * g.acc_scale * l2_10);
            dst.y = static_cast<float>(static_cast<double>(g.p1[n]) * l2_10);
            dst.z = (b >= 0 && b == first) ? weight[n] : 0.f;
            dst.w = (b >= 0 && b == first + 1) ? weight[n] : 0.f;
        }
    }
    DeviceGuard guard(e->device);
    if (g.depi_b) { cudaDeviceSynchronize(); cudaFree(g.depi_b); g.depi_b = nullptr; }
    return upload(&g.depi_b, epi.data(), epi.size());
}

int cpb200_engine_status(cpb200_engine* e, int32_t* watchdog_code, int32_t* range_flag, int clear_range) {
    if (!e) return fail(CPB200_E_INVALID, "null engine");
    volatile int* st = reinterpret_cast<volatile int*>(e->h_status);
    if (watchdog_code) *watchdog_code = st ? st[0] : 0;
    if (range_flag) *range_flag = st ? st[1] : 0;
    if (st && clear_range) st[1] = 0;
    if (st && st[0]) return fail(CPB200_E_CUDA, "fused-kernel watchdog code %d: %s", st[0], watchdog_text(st[0]));
    return 0;
}

int64_t cpb200_engine_last_kernel_cycles(cpb200_engine* e) {
    if (!e || !e->d_max_cycles) return -1;
    DeviceGuard guard(e->device);
    unsigned long long v = 0;
    if (!e->timing || cudaDeviceSynchronize() != cudaSuccess ||
        cudaMemcpy(&v, e->d_max_cycles, sizeof v, cudaMemcpyDeviceToHost) != cudaSuccess) return -1;
    return static_cast<int64_t>(v);
}

int cpb200_engine_set_timing(cpb200_engine* e, int enabled) {
    if (!e) return fail(CPB200_E_INVALID, "null engine");
    e->timing = enabled != 0;
    e->ev_valid = false;
    return 0;
}

float cpb200_engine_last_kernel_ms(cpb200_engine* e) {
    if (!e || !e->ev_valid) return -1.f;
    DeviceGuard guard(e->device);
    float ms = -1.f;
    if (cudaEventSynchronize(e->ev1) != cudaSuccess || cudaEventElapsedTime(&ms, e->ev0, e->ev1) != cudaSuccess) {
        cudaGetLastError();
        return -1.f;
    }
    return ms;
}

}  // extern "C"


// ---------------------------------------------------------------------------------------------
// Planck-lite likelihood tail
// ---------------------------------------------------------------------------------------------
struct cpb200_plik {
    int device = 0;
    int n_ell = 0, n_bins = 0, n_gather = 0, pad = 0, n_jt = 0, num_sms = 0;
    int in_len[3] = {0, 0, 0};          // floats per input row: n_ell, or the TE bin count when TE arrives binned
    bool te_binned = false;
    bool partials = false;              // TT and EE arrive as per-4-multipole partial band powers
    float units_factor = 0.f;
    cpb::PlikSlot* d_slots = nullptr;   // runs of <= PLIK_SEG consecutive multipoles of one bin (plik_bin_kernel)
    int n_passes = 0;
    size_t bin_smem = 0;
    float* d_fprime = nullptr;          // pad x pad: strict upper triangle of F + half its diagonal, zero elsewhere
    float* d_diff = nullptr;            // chunk_rows x pad
    float* d_partial = nullptr;         // chunk_rows x n_jt (fp32 path) or x 2 n_chunks (tensor path)
    // tensor-core quadratic form: the fused kernel with one upper-triangular contraction and the EPI_QUADFORM epilogue
    cpb200_engine* qf = nullptr;
    uint8_t* d_tiles = nullptr;         // split-fp16 operand tiles of one row chunk (written by plik_bin_kernel)
    float* d_row_scale = nullptr;       // chunk_rows
    int n_part = 0;
    int64_t chunk_rows = 0;
    int64_t launches = 0;
};

namespace {
void plik_free(cpb200_plik* p) {
    if (!p) return;
    cudaFree(p->d_slots);
    cudaFree(p->d_fprime); cudaFree(p->d_diff); cudaFree(p->d_partial); cudaFree(p->d_tiles); cudaFree(p->d_row_scale);
    if (p->qf) cpb200_destroy(p->qf);
    delete p;
}
}  // namespace

extern "C" {

int cpb200_plik_create(const cpb200_plik_desc* d, int32_t device, cpb200_plik** out) {
    if (!d || !out) return fail(CPB200_E_INVALID, "null argument");
    *out = nullptr;
    if (d->n_ell < 1 || d->n_bins < 1 || d->n_gather < 1 || !d->bin_start || !d->gather_first || !d->window ||
        !d->x_data || !d->fisher)
        return fail(CPB200_E_INVALID, "incomplete likelihood description");
    if (d->bin_start[0] != 0 || d->bin_start[d->n_bins] != d->n_gather) return fail(CPB200_E_INVALID, "bin_start does not span n_gather");
    for (int b = 0; b < d->n_bins; ++b) {
        const int len = d->bin_start[b + 1] - d->bin_start[b];
        if (len < 0 || d->gather_first[b] < 0 || d->gather_first[b] + len > 3 * d->n_ell)
            return fail(CPB200_E_INVALID, "bin %d gathers outside the three spectra", b);
        if (d->gather_first[b] % d->n_ell + len > d->n_ell)
            return fail(CPB200_E_INVALID, "bin %d straddles two spectra (multipoles %d .. %d of a %d-multipole row)", b,
                        d->gather_first[b] % d->n_ell, d->gather_first[b] % d->n_ell + len - 1, d->n_ell);
    }
    if (d->te_binned) {        // the TE band powers arrive as one contiguous run: its bins must be contiguous in the vector
        int first = -1, last = -1, count = 0;
        for (int b = 0; b < d->n_bins; ++b)
            if (d->gather_first[b] / d->n_ell == 1) { if (first < 0) first = b; last = b; ++count; }
        if (count && last - first + 1 != count) return fail(CPB200_E_INVALID, "te_binned needs the TE bins to be contiguous (bins %d .. %d hold %d)", first, last, count);
    }
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        return fail(CPB200_E_NO_DEVICE, "no CUDA device visible; this engine has no CPU fallback");
    }
    if (device < 0 || device >= ndev) return fail(CPB200_E_INVALID, "device %d out of range (%d visible)", device, ndev);
    DeviceGuard guard(device);
    cudaDeviceProp prop;
    CU(cudaGetDeviceProperties(&prop, device));
    // split the bins into slots of at most PLIK_SEG multipoles (plik_bin_kernel); a bin's slots sit on neighbouring
    // lanes of one warp (dummy slots pad to the next warp when they would straddle it)
    const bool te_binned = d->te_binned != 0;
    // which spectrum a bin belongs to, and the bins of TE (needed when TE arrives binned)
    int te_bin0 = -1, te_bins = 0;
    for (int b = 0; b < d->n_bins; ++b)
        if (d->gather_first[b] / d->n_ell == 1) { if (te_bin0 < 0) te_bin0 = b; ++te_bins; }
    if (te_binned && te_bins == 0) return fail(CPB200_E_INVALID, "te_binned set, but no bin gathers from the TE spectrum");
    const bool partials = d->partials != 0;
    const int part_len = 2 * ((d->n_ell + 3) / 4);
    const int in_len[3] = {partials ? part_len : d->n_ell, te_binned ? te_bins : d->n_ell, partials ? part_len : d->n_ell};
    // partial layout: which of a 4-multipole group's two partials belongs to bin b = is b the first bin the group touches
    std::vector<int> first_bin_of_group;
    if (partials) {
        first_bin_of_group.assign(static_cast<size_t>(3) * ((d->n_ell + 3) / 4), -1);
        for (int b = 0; b < d->n_bins; ++b) {
            const int sp = d->gather_first[b] / d->n_ell, lo = d->gather_first[b] % d->n_ell, len = d->bin_start[b + 1] - d->bin_start[b];
            for (int l = lo; l < lo + len; ++l) {
                int& fb = first_bin_of_group[static_cast<size_t>(sp) * ((d->n_ell + 3) / 4) + l / 4];
                if (fb < 0) fb = b;
            }
        }
    }
    const int in_off[3] = {0, round_up(in_len[0], 4), round_up(in_len[0], 4) + round_up(in_len[1], 4)};
    const int row_floats = in_off[2] + round_up(in_len[2], 4);
    std::vector<cpb::PlikSlot> slots;
    auto dummy = [] { cpb::PlikSlot s{}; s.bin = -1; return s; };
    for (int b = 0; b < d->n_bins; ++b) {
        const int sp = d->gather_first[b] / d->n_ell, lo = d->gather_first[b] % d->n_ell;
        const float sqrt_f = std::sqrt(std::max(0.f, d->fisher[static_cast<size_t>(b) * d->n_bins + b]));
        if (sp == 1 && te_binned) {          // one element, already the band power
            cpb::PlikSlot s = dummy();
            s.off = in_off[1] + (b - te_bin0); s.last = 0; s.w[0] = 1.f;
            s.bin = b; s.n_slots = 1; s.x_data = d->x_data[b]; s.sqrt_f = sqrt_f;
            slots.push_back(s);
            continue;
        }
        const int len = d->bin_start[b + 1] - d->bin_start[b];
        if (partials && sp != 1) {
            // the bin's partials: one per 4-multipole group it touches, at 2 * group + (0: first bin of the group, 1: second);
            // covered by slots of PLIK_SEG consecutive elements with weight 1 on the bin's own partials
            std::vector<int> pos;
            const int ng = (d->n_ell + 3) / 4;
            for (int gq = lo / 4; gq <= (lo + len - 1) / 4; ++gq) {
                const int fb = first_bin_of_group[static_cast<size_t>(sp) * ng + gq];
                if (fb != b && fb + 1 != b) return fail(CPB200_E_UNSUPPORTED, "multipoles %d .. %d touch more than two bins", 4 * gq, 4 * gq + 3);
                pos.push_back(2 * gq + (fb == b ? 0 : 1));
            }
            std::vector<cpb::PlikSlot> mine;
            for (size_t i = 0; i < pos.size();) {
                cpb::PlikSlot s = dummy();
                const int p0 = pos[i];
                s.off = in_off[sp] + p0;
                while (i < pos.size() && pos[i] - p0 < cpb::PLIK_SEG) { s.w[pos[i] - p0] = 1.f; s.last = pos[i] - p0; ++i; }
                mine.push_back(s);
            }
            const int nsl = static_cast<int>(mine.size());
            if (nsl < 1 || nsl > cpb::PLIK_MAX_SLOTS_PER_BIN)
                return fail(CPB200_E_UNSUPPORTED, "bin %d spans %d partial slots; 1 .. %d are supported", b, nsl, cpb::PLIK_MAX_SLOTS_PER_BIN);
            while (static_cast<int>(slots.size() % 32) + nsl > 32) slots.push_back(dummy());
            mine[0].bin = b; mine[0].n_slots = nsl; mine[0].x_data = d->x_data[b]; mine[0].sqrt_f = sqrt_f;
            for (const cpb::PlikSlot& s : mine) slots.push_back(s);
            continue;
        }
        const int ns = (len + cpb::PLIK_SEG - 1) / cpb::PLIK_SEG;
        if (ns < 1 || ns > cpb::PLIK_MAX_SLOTS_PER_BIN)
            return fail(CPB200_E_UNSUPPORTED, "bin %d gathers %d multipoles; 1 .. %d are supported", b, len, cpb::PLIK_SEG * cpb::PLIK_MAX_SLOTS_PER_BIN);
        while (static_cast<int>(slots.size() % 32) + ns > 32) slots.push_back(dummy());
        for (int k = 0; k < ns; ++k) {
            cpb::PlikSlot s = dummy();
            const int j0 = d->bin_start[b] + k * cpb::PLIK_SEG, j1 = std::min(j0 + cpb::PLIK_SEG, d->bin_start[b + 1]);
            s.off = in_off[sp] + lo + k * cpb::PLIK_SEG;
            s.last = j1 - j0 - 1;
            for (int u = 0; u < j1 - j0; ++u) s.w[u] = d->window[j0 + u];
            if (k == 0) { s.bin = b; s.n_slots = ns; s.x_data = d->x_data[b]; s.sqrt_f = sqrt_f; }
            slots.push_back(s);
        }
    }
    while (slots.size() % cpb::PLIK_BIN_THREADS) slots.push_back(dummy());
    const int n_passes = static_cast<int>(slots.size() / cpb::PLIK_BIN_THREADS);
    if (n_passes > cpb::PLIK_MAX_PASSES)
        return fail(CPB200_E_UNSUPPORTED, "%zu window slots exceed %d", slots.size(), cpb::PLIK_MAX_PASSES * cpb::PLIK_BIN_THREADS);
    const size_t bin_smem = (cpb::PLIK_BAR_FLOATS + static_cast<size_t>(cpb::PLIK_GROUPS * cpb::PLIK_R) * (row_floats + cpb::PLIK_ROW_PAD) +
                             2 * cpb::PLIK_R * (round_up(d->n_bins, cpb::PLIK_QF_BN) + cpb::PLIK_BIN_THREADS / 32)) * sizeof(float);
    if (bin_smem > prop.sharedMemPerBlockOptin)
        return fail(CPB200_E_UNSUPPORTED, "the row ring of 4 x 3 x %d multipoles does not fit one block's shared memory", d->n_ell);
    cpb200_plik* p = new cpb200_plik();
    auto bail = [&](int rc) { plik_free(p); return rc; };
    p->device = device;
    p->num_sms = prop.multiProcessorCount;
    p->n_ell = d->n_ell; p->n_bins = d->n_bins; p->n_gather = d->n_gather;
    p->te_binned = te_binned;
    p->partials = partials;
    for (int i = 0; i < 3; ++i) p->in_len[i] = in_len[i];
    p->units_factor = d->units_factor;
    p->pad = round_up(d->n_bins, cpb::PLIK_QF_BN);
    p->n_jt = p->pad / cpb::PLIK_QF_BN;
    int rc;
    p->n_passes = n_passes;
    p->bin_smem = bin_smem;
    if ((rc = upload(&p->d_slots, slots.data(), slots.size()))) return bail(rc);
    if (getenv("CPB200_PLIK_FP32") != nullptr) {   // fp32 cross-check path: F' = strict upper triangle of the symmetrised F
        // + half its diagonal, chi^2 = 2 d^T F' d
        std::vector<float> fp(static_cast<size_t>(p->pad) * p->pad, 0.f);
        for (int i = 0; i < d->n_bins; ++i)
            for (int j = i; j < d->n_bins; ++j) {
                const double fij = 0.5 * (static_cast<double>(d->fisher[static_cast<size_t>(i) * d->n_bins + j]) +
                                          static_cast<double>(d->fisher[static_cast<size_t>(j) * d->n_bins + i]));
                if (!std::isfinite(fij)) return bail(fail(CPB200_E_RANGE, "non-finite Fisher element"));
                fp[static_cast<size_t>(i) * p->pad + j] = static_cast<float>(i == j ? 0.5 * fij : fij);
            }
        if ((rc = upload(&p->d_fprime, fp.data(), fp.size()))) return bail(rc);
    }
    p->chunk_rows = 131072;
    if (getenv("CPB200_PLIK_FP32") == nullptr) {
        // Tensor-core quadratic form.  With s_i = sqrt(F_ii): chi^2 = 2 v^T G v, v_i = d_i s_i, G = triu(F,1)/(s_i s_j) +
        // I/2 - a correlation-like matrix with entries in [-1, 1], well inside the split-fp16 range; the rows of v are
        // brought to [0.5, 1) by a power of two in plik_bin_kernel.  The contraction runs in the fused kernel
        // (one upper-triangular 640 x 640 contraction, EPI_QUADFORM epilogue) on operand tiles the binning kernel prebuilds.
        cpb200_engine* q = new cpb200_engine();
        p->qf = q;
        q->device = device; q->precision = CPB200_PREC_F16X3; q->P = p->pad; q->n_modes = p->pad; q->num_sms = p->num_sms;
        q->mP[0] = p->pad; q->mModes[0] = p->pad;
        Gemm g;
        g.K = p->pad; g.N = p->pad; g.epi = cpb::EPI_QUADFORM; g.tri = true;
        g.W.assign(static_cast<size_t>(p->pad) * p->pad, 0.f);
        g.p0.assign(p->pad, 0.f); g.p1.assign(p->pad, 0.f); g.p2.assign(p->pad, 0.f);
        for (int i = 0; i < d->n_bins; ++i) {
            const double si = std::sqrt(static_cast<double>(d->fisher[static_cast<size_t>(i) * d->n_bins + i]));
            if (!(si > 0.0) || !std::isfinite(si)) return bail(fail(CPB200_E_RANGE, "Fisher diagonal element %d is not positive", i));
            for (int j = i; j < d->n_bins; ++j) {
                const double sj = std::sqrt(static_cast<double>(d->fisher[static_cast<size_t>(j) * d->n_bins + j]));
                const double fij = 0.5 * (static_cast<double>(d->fisher[static_cast<size_t>(i) * d->n_bins + j]) +
                                          static_cast<double>(d->fisher[static_cast<size_t>(j) * d->n_bins + i]));
                g.W[static_cast<size_t>(i) * p->pad + j] = static_cast<float>(i == j ? 0.5 : fij / (si * sj));
            }
        }
        q->gemms.push_back(std::move(g));
        if ((rc = setup_umma(q))) return bail(rc);
        for (auto& gg : q->gemms) { gg.W.clear(); gg.W.shrink_to_fit(); }
        if (cudaEventCreate(&q->ev0) != cudaSuccess || cudaEventCreate(&q->ev1) != cudaSuccess)
            return bail(fail(CPB200_E_CUDA, "event creation failed"));
        p->n_part = 2 * q->gemms[0].n_chunks;
        const size_t n_tiles = static_cast<size_t>(round_up(static_cast<int>((p->chunk_rows + 127) / 128), 4));
        if (cudaMalloc(reinterpret_cast<void**>(&p->d_tiles), n_tiles * q->um_a0_bytes) != cudaSuccess ||
            cudaMemset(p->d_tiles, 0, n_tiles * q->um_a0_bytes) != cudaSuccess ||
            cudaMalloc(reinterpret_cast<void**>(&p->d_row_scale), static_cast<size_t>(p->chunk_rows) * sizeof(float)) != cudaSuccess)
            return bail(fail(CPB200_E_CUDA, "likelihood tile allocation failed: %s", cudaGetErrorString(cudaGetLastError())));
    }
    const int partial_cols = std::max(p->n_jt, p->n_part);
    if (cudaMalloc(reinterpret_cast<void**>(&p->d_diff), static_cast<size_t>(p->chunk_rows) * p->pad * sizeof(float)) != cudaSuccess ||
        cudaMalloc(reinterpret_cast<void**>(&p->d_partial), static_cast<size_t>(p->chunk_rows) * partial_cols * sizeof(float)) != cudaSuccess)
        return bail(fail(CPB200_E_CUDA, "likelihood scratch allocation failed: %s", cudaGetErrorString(cudaGetLastError())));
    // the padding columns n_bins .. pad-1 of the difference vectors are never written: zero them once
    if (cudaMemset(p->d_diff, 0, static_cast<size_t>(p->chunk_rows) * p->pad * sizeof(float)) != cudaSuccess)
        return bail(fail(CPB200_E_CUDA, "memset failed"));
    if (cudaFuncSetAttribute(cpb::plik_bin_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(bin_smem)) != cudaSuccess ||
        cudaFuncSetAttribute(cpb::plik_bin_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(bin_smem)) != cudaSuccess)
        return bail(fail(CPB200_E_CUDA, "cannot size the binning kernel's shared memory"));
    *out = p;
    return 0;
}

int cpb200_plik_destroy(cpb200_plik* p) {
    if (!p) return 0;
    DeviceGuard guard(p->device);
    cudaDeviceSynchronize();
    plik_free(p);
    return 0;
}

int cpb200_plik_loglike_device(cpb200_plik* p, const float* cl_tt, const float* cl_te, const float* cl_ee,
                               int64_t pitch, int64_t te_pitch, const float* cal, int64_t n_rows, float* loglike,
                               float* binned, int64_t binned_pitch, void* stream) {
    if (!p) return fail(CPB200_E_INVALID, "null likelihood");
    if (n_rows < 0) return fail(CPB200_E_INVALID, "negative row count");
    if (n_rows == 0) return 0;
    if (!cl_tt || !cl_te || !cl_ee || !cal || !loglike) return fail(CPB200_E_INVALID, "null device pointer");
    if (pitch < p->in_len[0]) return fail(CPB200_E_INVALID, "pitch %lld < %d floats per TT / EE row", static_cast<long long>(pitch), p->in_len[0]);
    if (te_pitch < p->in_len[1]) return fail(CPB200_E_INVALID, "te_pitch %lld < %d", static_cast<long long>(te_pitch), p->in_len[1]);
    if (binned && binned_pitch < p->n_bins) return fail(CPB200_E_INVALID, "binned_pitch < n_bins");
    DeviceGuard guard(p->device);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const size_t smem = p->bin_smem;
    for (int64_t r0 = 0; r0 < n_rows; r0 += p->chunk_rows) {
        const int64_t m = std::min<int64_t>(p->chunk_rows, n_rows - r0);
        const int grid_bin = static_cast<int>(std::min<int64_t>(m, static_cast<int64_t>(p->num_sms) * cpb::PLIK_BLOCKS_PER_SM));
        cpb::PlikInputs in;
        in.ptr[0] = cl_tt + r0 * pitch; in.ptr[1] = cl_te + r0 * te_pitch; in.ptr[2] = cl_ee + r0 * pitch;
        in.pitch[0] = pitch; in.pitch[1] = te_pitch; in.pitch[2] = pitch;
        for (int i = 0; i < 3; ++i) in.len[i] = p->in_len[i];
        if (p->qf)
            cpb::plik_bin_kernel<true><<<grid_bin, cpb::PLIK_BIN_THREADS, smem, st>>>(
                in, cal + r0, m, p->d_slots, p->n_passes, p->units_factor, p->d_diff, p->pad,
                binned ? binned + r0 * binned_pitch : nullptr, binned_pitch, p->d_tiles, p->d_row_scale);
        else
            cpb::plik_bin_kernel<false><<<grid_bin, cpb::PLIK_BIN_THREADS, smem, st>>>(
                in, cal + r0, m, p->d_slots, p->n_passes, p->units_factor, p->d_diff, p->pad,
                binned ? binned + r0 * binned_pitch : nullptr, binned_pitch, nullptr, nullptr);
        if (p->qf) {
            cpb200_engine* q = p->qf;
            q->qf_ext_a0 = p->d_tiles; q->qf_d = p->d_diff; q->qf_pitch = p->pad; q->qf_cols = p->pad; q->qf_partial = p->d_partial;
            if (int rc = order_before(q, st)) return rc;
            if (int rc = launch_umma(q, nullptr, m, nullptr, nullptr, nullptr, st)) return rc;
            if (int rc = order_after(q, st)) return rc;
            cpb::plik_finish_tensor_kernel<<<static_cast<unsigned>((m + 255) / 256), 256, 0, st>>>(
                p->d_partial, p->n_part, p->d_row_scale, q->gemms[0].acc_scale, m, loglike + r0);
        } else {
            dim3 grid_qf(static_cast<unsigned>(p->n_jt), static_cast<unsigned>((m + cpb::PLIK_QF_BM - 1) / cpb::PLIK_QF_BM));
            cpb::plik_quadform_kernel<<<grid_qf, 256, 0, st>>>(p->d_diff, p->pad, p->d_fprime, p->pad, m, p->d_partial, p->n_jt);
            cpb::plik_finish_kernel<<<static_cast<unsigned>((m + 255) / 256), 256, 0, st>>>(p->d_partial, p->n_jt, m, loglike + r0);
        }
        p->launches += 3;
    }
    CU(cudaGetLastError());
    return 0;
}

int64_t cpb200_plik_launch_count(const cpb200_plik* p) { return p ? p->launches : 0; }

}  // extern "C"
