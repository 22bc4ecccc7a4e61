// Microbenchmark: the fused kernel's epilogues in isolation - eight epilogue warps of a 384-thread CTA per SM drain
// 128 x 256 fp32 accumulator chunks from TMEM (whatever it holds), run the real epilogue code of umma_kernel.cuh
// (activation + fp16 split + operand-tile stores, or de-standardise + 10**y + spectrum stores) and write to an
// L2-resident scratch / an output matrix.  No loader, no MMA: what is measured is the epilogue's own rate (cycles per
// chunk per SM), the figure the pipeline's epilogue-bound phases run at.
// nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -I../../cosmopower_b200/csrc -o epi_pipe epi_pipe.cu
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "umma_kernel.cuh"

using namespace cpb;

// KIND 0: activation chunks (EPI_ACT), 1: output chunks with 10**y (EPI_OUT fast path)
template <int KIND, int DBG, int SPIN>
__global__ void __launch_bounds__(UM_THREADS, 1) epi_bench(uint8_t* scratch, float* out, long long out_pitch, const float4* consts, int jobs,
                                                           long long* cyc, int nc, int pace, const uint8_t* wsrc, int lmode, int nreg)
{
    extern __shared__ uint8_t smem_raw[];
    __shared__ uint32_t tmem_ptr;
    __shared__ unsigned long long spin_bar;
    __shared__ volatile int done;
    float4* par = reinterpret_cast<float4*>(smem_raw);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < UM_MAX_NC; i += blockDim.x) par[i] = consts[i];
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_ptr)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (threadIdx.x == 0) { mbar_init(smem_u32(&spin_bar), 1); done = 0; }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_ptr;
    if (SPIN == 4 && warp == 0) {
        // the pipeline's loader: two 16 KB bulk copies per stage (an activation stage of this CTA's scratch, a weight stage)
        // into a six-deep ring, one stage every `pace` cycles
        __shared__ unsigned long long full[UM_STAGES];
        uint8_t* ring = smem_raw + UM_MAX_NC * 16;
        if (lane == 0) for (int i = 0; i < UM_STAGES; ++i) mbar_init(smem_u32(&full[i]), 1);
        __syncwarp();
        const uint8_t* asrc = scratch + static_cast<size_t>(blockIdx.x) * nreg * (1u << 17);
        int stage = 0; long long n = 0; long long tn = clock64();
        while (!done) {
            if (n >= UM_STAGES) { while (!mbar_try_wait(smem_u32(&full[stage]), ((n / UM_STAGES) - 1) & 1)) {} }
            while (clock64() < tn) __nanosleep(64);
            tn += pace;
            if (lane == 0) {
                const uint32_t dst = smem_u32(ring + stage * UM_STAGE_BYTES);
                mbar_arrive_expect_tx(smem_u32(&full[stage]), 32768);
                const uint8_t* a = asrc + (n % (8 * nreg)) * 16384;
                const uint8_t* w = wsrc + (n % 256) * 16384 + (blockIdx.x & 1) * 8192;
                // lmode 0: activation stage + weight stage (the kernel's), 1: weights only (no line of the scratch is read),
                // 2: scratch only, 3: a private read-only region instead of the scratch
                if (lmode == 1) a = wsrc + ((n + 128) % 256) * 16384;
                if (lmode == 2) w = asrc + ((n + 4 * nreg) % (8 * nreg)) * 16384;
                if (lmode == 3) a = wsrc + (8u << 20) + static_cast<size_t>(blockIdx.x) * (3u << 18) + (n % 48) * 16384;
                bulk_g2s(dst, a, 16384, smem_u32(&full[stage]));
                bulk_g2s(dst + 16384, w, 16384, smem_u32(&full[stage]));
            }
            __syncwarp();
            ++n;
            if (++stage == UM_STAGES) stage = 0;
        }
        for (long long m = (n > UM_STAGES ? n - UM_STAGES : 0); m < n; ++m)       // drain the copies still in flight
            while (!mbar_try_wait(smem_u32(&full[m % UM_STAGES]), (m / UM_STAGES) & 1)) {}
    } else
    if (SPIN && SPIN < 4 && warp < SPIN) {
        // the pipeline's other roles (loader, MMA issuer, standardiser) mostly wait on mbarriers while the epilogue is the bound:
        // the same wait loop as the kernel's, on a barrier that never completes
        while (!done) {
            if (mbar_try_wait(smem_u32(&spin_bar), 0)) break;
            if (clock64() == 0) break;
        }
    }
    long long t0 = clock64();
    if (warp >= UM_EPI_WARP0) {
        EpiCtx E;
        E.q = warp & 3; E.cw = (warp - UM_EPI_WARP0) >> 2; E.tq = lane >> 2; E.tr = lane & 3; E.lane = lane;
        E.dbg = DBG; E.leader = true; E.rel_bar = 0; E.nc = nc ? nc : 256; E.par = par;
        uint8_t* cta_scratch = scratch + static_cast<size_t>(blockIdx.x) * nreg * (1u << 17);
        for (int j = 0; j < jobs; ++j) {
            const int slot = (j & 1) * 2;
            E.t_addr0 = tmem_base + (static_cast<uint32_t>(E.q * 32) << 16) + static_cast<uint32_t>(slot * UM_SLOT_COLS);
            E.chunk_col0 = 0;
            if (KIND == 0) {
                epi_chunk_operand<EPI_ACT>(E, 1.f / 1024.f, cta_scratch + static_cast<size_t>(j % nreg) * (1u << 17), nullptr);   // a job writes 8 stages x 16 KB
            } else {
                // 10 chunks of 256 columns per 128-row tile; tiles walk down the output matrix
                const long long row0 = (static_cast<long long>(blockIdx.x) * ((jobs + 9) / 10) + j / 10) * UM_TILE_M;
                E.chunk_col0 = (j % 10) * 256;
                float* orow0 = out + (row0 + E.q * 32 + E.tq) * out_pitch;
                epi_chunk_output<true, false>(E, orow0, out_pitch, OUT_VEC32, 1 << 30, 2560);
            }
            tc_fence_before();
            __syncwarp();
        }
    }
    long long t1 = clock64();
    if (threadIdx.x == UM_EPI_WARP0 * 32) cyc[blockIdx.x] = t1 - t0;
    if (warp >= UM_EPI_WARP0) { __syncwarp(); if (warp == UM_EPI_WARP0 + UM_EPI_WARPS - 1 && lane == 0) done = 1; }
    __syncthreads();
    if (warp == 2) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
}

int main(int argc, char** argv) {
    const int jobs = argc > 1 ? atoi(argv[1]) : 400;
    const int nc_arg = argc > 2 ? atoi(argv[2]) : 0;      // 0: chunk width known at compile time (256), else passed at run time
    const int blocks = 148;
    uint8_t* scratch; float* out; float4* consts; long long* cyc;
    cudaMalloc(&scratch, static_cast<size_t>(blocks) * (3u << 18));
    const long long pitch = 2560, rows = static_cast<long long>(blocks) * ((jobs + 9) / 10) * UM_TILE_M;
    cudaMalloc(&out, rows * pitch * sizeof(float));
    cudaMalloc(&cyc, blocks * sizeof(long long));
    std::vector<float4> h(UM_MAX_NC);
    for (int i = 0; i < UM_MAX_NC; ++i) h[i] = (i & 4) ? make_float4(8.f, 8.f, 56.f, 56.f) : make_float4(0.1f, 0.2f, -1.3f, -0.7f);
    cudaMalloc(&consts, h.size() * sizeof(float4));
    cudaMemcpy(consts, h.data(), h.size() * sizeof(float4), cudaMemcpyHostToDevice);
    const int smem = UM_MAX_NC * 16 + UM_STAGES * UM_STAGE_BYTES + 1024;
    const int pace = argc > 3 ? atoi(argv[3]) : 770;
    uint8_t* wsrc; cudaMalloc(&wsrc, (8u << 20) + 148u * (3u << 18)); cudaMemset(wsrc, 0, (8u << 20) + 148u * (3u << 18));
    const int lmode = argc > 4 ? atoi(argv[4]) : 0;
    const int nreg = argc > 5 ? atoi(argv[5]) : 6;       // CTA-private scratch in units of 128 KB (the kernel's: 6)
    cudaFuncSetAttribute(epi_bench<0, 0, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaFuncSetAttribute(epi_bench<1, 0, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaFuncSetAttribute(epi_bench<0, 0, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaFuncSetAttribute(epi_bench<1, 0, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    for (int kind = 0; kind < 2; ++kind) {
        for (int dbg : {0, 4}) {          // here: number of waiting role warps
            for (int rep = 0; rep < 2; ++rep) {
                if (kind == 0 && dbg == 0) epi_bench<0, 0, 0><<<blocks, UM_THREADS, smem>>>(scratch, out, pitch, consts, jobs, cyc, nc_arg, pace, wsrc, lmode, nreg);
                if (kind == 0 && dbg == 1) epi_bench<0, 0, 1><<<blocks, UM_THREADS, smem>>>(scratch, out, pitch, consts, jobs, cyc, nc_arg, pace, wsrc, lmode, nreg);
                if (kind == 0 && dbg == 2) epi_bench<0, 0, 2><<<blocks, UM_THREADS, smem>>>(scratch, out, pitch, consts, jobs, cyc, nc_arg, pace, wsrc, lmode, nreg);
                if (kind == 0 && dbg == 4) epi_bench<0, 0, 4><<<blocks, UM_THREADS, smem>>>(scratch, out, pitch, consts, jobs, cyc, nc_arg, pace, wsrc, lmode, nreg);
                if (kind == 1 && dbg == 4) epi_bench<1, 0, 4><<<blocks, UM_THREADS, smem>>>(scratch, out, pitch, consts, jobs, cyc, nc_arg, pace, wsrc, lmode, nreg);
                if (kind == 0 && dbg == 3) epi_bench<0, 0, 3><<<blocks, UM_THREADS, smem>>>(scratch, out, pitch, consts, jobs, cyc, nc_arg, pace, wsrc, lmode, nreg);
                if (kind == 1 && dbg == 0) epi_bench<1, 0, 0><<<blocks, UM_THREADS, smem>>>(scratch, out, pitch, consts, jobs, cyc, nc_arg, pace, wsrc, lmode, nreg);
                if (kind == 1 && dbg == 1) epi_bench<1, 0, 1><<<blocks, UM_THREADS, smem>>>(scratch, out, pitch, consts, jobs, cyc, nc_arg, pace, wsrc, lmode, nreg);
                if (kind == 1 && dbg == 2) epi_bench<1, 0, 2><<<blocks, UM_THREADS, smem>>>(scratch, out, pitch, consts, jobs, cyc, nc_arg, pace, wsrc, lmode, nreg);
                if (kind == 1 && dbg == 3) epi_bench<1, 0, 3><<<blocks, UM_THREADS, smem>>>(scratch, out, pitch, consts, jobs, cyc, nc_arg, pace, wsrc, lmode, nreg);
            }
            if (cudaDeviceSynchronize() != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(cudaGetLastError())); return 1; }
            std::vector<long long> c(blocks);
            cudaMemcpy(c.data(), cyc, blocks * sizeof(long long), cudaMemcpyDeviceToHost);
            double s = 0, mx = 0;
            for (long long v : c) { s += v; mx = mx > v ? mx : v; }
            printf("%s, %d role warps waiting (4 = loader streaming stages): %.0f cycles per 128x256 chunk per SM (mean over SMs; slowest %.0f)\n",
                   kind == 0 ? "activation + split + operand stores" : "de-standardise + 10**y + spectrum stores", dbg, s / blocks / jobs, mx / jobs);
        }
    }
    return 0;
}
