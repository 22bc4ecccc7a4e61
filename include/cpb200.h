/*
 * cpb200.h — C ABI of the B200 engine for CosmoPower's batched emulator forward pass.
 *
 * The reference (alessiospuriomancini/cosmopower) is pure Python and has no FFI layer: its
 * boundary for this path is the Python class API.  Each entry point below names the reference
 * interface it stands in for; cosmopower_b200/cosmopower_NN.py and cosmopower_PCAplusNN.py bind
 * these symbols with ctypes and expose the reference's method names on top of them
 * (see INTEGRATION.md for the stub a reference maintainer would add).
 *
 * Conventions: plain pointers and sizes only; every function returns 0 on success or a negative
 * CPB200_E_* code and never throws or aborts; cpb200_last_error() gives the message of the last
 * failure on the calling thread.  One engine per (model, device).  Calls on different engines
 * may run concurrently.  Launches of one engine share its activation scratch: the library orders a
 * launch on another stream behind the engine's previous launch with an event, so asynchronous calls
 * on different streams are safe; calls from several host THREADS on one engine must still be
 * serialised by the caller (the Python binding holds a lock per engine).
 */
#ifndef CPB200_H
#define CPB200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CPB200_ABI_VERSION 3

/* error codes */
#define CPB200_OK            0
#define CPB200_E_INVALID    -1   /* bad argument / inconsistent model description        */
#define CPB200_E_CUDA       -2   /* a CUDA runtime call failed (message has the detail)  */
#define CPB200_E_NO_DEVICE  -3   /* no sm_100 device visible: there is NO CPU fallback   */
#define CPB200_E_UNSUPPORTED -4  /* layer shape outside what the kernels were built for  */
#define CPB200_E_RANGE      -5   /* weights, or a batch's inputs / activations, outside the fp16 split range
                                    (use CPB200_PREC_FP32_SIMT) */

/* numeric formats of the dense contractions (one kernel family, not back-ends) */
#define CPB200_PREC_FP32_SIMT  0  /* CUDA-core fp32 FMA: on-GPU fp32 truth, slow               */
#define CPB200_PREC_F16X3      1  /* tcgen05 kind::f16, 2-way fp16 split, 3 products, fp32 acc */

/* flags for cpb200_forward_* */
#define CPB200_FLAG_TEN_TO     1  /* apply 10**x to the output (ten_to_predictions_np)         */
#define CPB200_FLAG_ACCUMULATE 2  /* device path only: add the log10 values already in out_dev before the optional
                                     10**x - fuses `10.**(lin.predictions_np(p) + boost.predictions_np(q))`
                                     (reference cosmopower/trained_models/CP_paper/PK/README.md:62-65)        */

#define CPB200_FLAG_BINNED     4  /* device path, with TEN_TO, after cpb200_engine_set_binning: write partial band powers of
                                     10**y instead of the spectrum (see cpb200_engine_set_binning)                          */

typedef struct cpb200_engine cpb200_engine;

/*
 * Host-side description of one restored emulator: exactly the attributes the reference's
 * `restore` sets (cosmopower/cosmopower_NN.py:310-327; cosmopower/cosmopower_PCAplusNN.py:306-326).
 * All arrays are float32, C-contiguous, owned by the caller and only read during cpb200_create.
 *   dims[0..n_dense]  = `architecture`   (n_parameters, hidden..., n_modes or n_pcas)
 *   W[i]              = `W_[i]`  (dims[i], dims[i+1]) row-major, applied as x @ W (no transpose)
 *   b[i]              = `b_[i]`  (dims[i+1])
 *   alphas[i]/betas[i]= `alphas_[i]` / `betas_[i]` (dims[i+1]),  i < n_dense-1
 * For cosmopower_PCAplusNN set n_pcas > 0 and the three pca_* arrays
 * (`pca_mean_`, `pca_std_` (n_pcas), `pca_transform_matrix_` (n_pcas, n_modes) row-major).
 */
typedef struct cpb200_model_desc {
    int32_t n_parameters;
    int32_t n_modes;
    int32_t n_dense;              /* == reference n_layers: number of W matrices */
    int32_t n_pcas;               /* 0 for cosmopower_NN */
    const int32_t* dims;
    const float* const* W;
    const float* const* b;
    const float* const* alphas;
    const float* const* betas;
    const float* parameters_mean;
    const float* parameters_std;
    const float* features_mean;
    const float* features_std;
    const float* pca_mean;
    const float* pca_std;
    const float* pca_transform_matrix;
} cpb200_model_desc;

/* Library identity: returns CPB200_ABI_VERSION. */
int cpb200_abi_version(void);

/* Number of visible CUDA devices with compute capability 10.x (0 if none / no driver). */
int cpb200_device_count(void);

/*
 * Replaces: constructor with restore=True followed by the TF variable set-up
 * (cosmopower_NN.py:46-68,92-118; cosmopower_PCAplusNN.py:37-54,80-126).
 * Uploads and packs the weights once on `device` (split into fp16 hi/lo, tiled for the tensor
 * cores).  `precision` is one of CPB200_PREC_*.
 */
int cpb200_create(const cpb200_model_desc* desc, int device, int precision, cpb200_engine** out);

/* Replaces: garbage collection of the reference object. Safe on NULL. */
int cpb200_destroy(cpb200_engine* e);

/*
 * Replaces: predictions_tf / ten_to_predictions_tf — array in, array out, both resident on the
 * device (cosmopower_NN.py:160-211; cosmopower_PCAplusNN.py:164-239), i.e. forward_pass_np
 * (cosmopower_NN.py:354-384; cosmopower_PCAplusNN.py:351-381) without the host round trip.
 *   params_dev : (n_rows, n_parameters) float32 row-major, already in the model's column order
 *   out_dev    : n_rows rows of n_modes float32, row pitch `out_pitch` elements (>= n_modes)
 *   stream     : cudaStream_t (NULL = legacy default stream); the call is asynchronous on it
 * No allocation happens inside when n_rows fits the engine's scratch (it always does: the engine
 * walks the batch in row tiles).  Caller owns both buffers.
 */
int cpb200_forward_device(cpb200_engine* e, const float* params_dev, int64_t n_rows,
                          float* out_dev, int64_t out_pitch, int flags, void* stream);

/*
 * Replaces: forward_pass_np / predictions_np / ten_to_predictions_np on host arrays
 * (cosmopower_NN.py:354-425; cosmopower_PCAplusNN.py:351-421).  Host float32 in, host float32
 * out; the copies host->device and device->host are pipelined with the kernel in row chunks.
 * Buffers from cpb200_host_alloc (pinned) are DMA'd directly; pageable buffers go through the
 * engine's pinned bounce buffers (a result of up to 2048 rows leaves in pieces behind ONE launch, each copied out while
 * the next are on the wire; larger ones in row chunks of an eighth of the batch).  The copy-out runs on a small pool of
 * library threads that are woken for the duration of a call with 4 MB of result or more and sleep otherwise
 * (CPB200_HOST_THREADS, default: the CPUs the process may run on, at most 32).  Synchronous: returns when `out_host` is complete.
 * Only the n_modes floats of each row are written: with out_pitch > n_modes the padding is left alone (pitched copy);
 * with out_pitch == n_modes a transfer of 256 MB or more leaves the device as one contiguous copy per chunk, which is
 * ~8 % faster over PCIe than the pitched copy - the layout to choose for bulk results.
 */
int cpb200_forward_host(cpb200_engine* e, const float* params_host, int64_t n_rows,
                        float* out_host, int64_t out_pitch, int flags);

/*
 * The same call with a float64 result: the reference computes in the dtype NumPy promotion gives its input
 * (cosmopower_NN.py:371-384, cosmopower_PCAplusNN.py:368-381: float64 for dicts of lists / float64 arrays), so
 * predictions_np returns float64 there.  The device result (float32) is widened while it is copied out of the pinned
 * bounce buffers, rows split over the library's host threads (see above); out_pitch in doubles.
 */
int cpb200_forward_host_f64(cpb200_engine* e, const float* params_host, int64_t n_rows,
                            double* out_host, int64_t out_pitch, int flags);

/*
 * Several emulators evaluated on the SAME rows by ONE launch.  Replaces the reference's back-to-back calls
 *   tt.ten_to_predictions_tf(x); te.predictions_tf(x); ee.ten_to_predictions_tf(x)
 *     (cosmopower/likelihoods/tf_planck2018_lite/tf_planck2018_lite.py:276-278: three emulators, one parameter tensor) and
 *   10.**(lin.predictions_np(p_lin) + boost.predictions_np(p_boost))
 *     (cosmopower/trained_models/CP_paper/PK/README.md:62-65: two emulators whose log-spectra are summed).
 * cpb200_create_multi packs up to 3 emulators (tensor-core format) into one engine; one persistent launch then walks every
 * pair of row tiles through all of them in turn - each emulator's first layer runs inside the previous emulator's output
 * phase, so the tensor pipe never waits for a layer-0 epilogue between emulators - and writes each emulator's spectra to
 * its own buffer.  params_dev[m] is emulator m's (n_rows, n_parameters[m]) array (the pointers may alias when the
 * emulators share their parameter columns); flags[m] as for cpb200_forward_device.  CPB200_FLAG_ACCUMULATE on emulator
 * m >= 1 adds its log10 result to what emulator m-1 has just written to the SAME buffer (same n_modes and pitch) before
 * the optional 10**: the sum is read back from L2 by the lane that wrote it, never from HBM.  Asynchronous on `stream`.
 */
int cpb200_create_multi(const cpb200_model_desc* const* descs, int32_t n_models, int device, cpb200_engine** out);
int cpb200_forward_multi_device(cpb200_engine* e, const float* const* params_dev, int64_t n_rows, float* const* out_dev,
                                const int64_t* out_pitch, const int32_t* flags, void* stream);

/*
 * The same two calls on several devices of one node from ONE process (rows are independent, every device
 * holds a full engine): `engines` lists one engine of the same model per device; rows are split into contiguous
 * shards (sizes differing by at most one row), one host thread per device - bound to the CPUs next to its GPU - runs the
 * pipelined host path of its engine and DMAs its rows straight into the caller's array.  No collective is involved.
 * Replaces: predictions_np / ten_to_predictions_np (cosmopower_NN.py:388-425; cosmopower_PCAplusNN.py:384-421) on a
 * multi-GPU box; the result is identical, bit for bit, to the single-device call.
 */
int cpb200_forward_host_multi(cpb200_engine* const* engines, int32_t n_engines, const float* params_host, int64_t n_rows,
                              float* out_host, int64_t out_pitch, int flags);
int cpb200_forward_host_multi_f64(cpb200_engine* const* engines, int32_t n_engines, const float* params_host, int64_t n_rows,
                                  double* out_host, int64_t out_pitch, int flags);

/*
 * Band-power binning fused into the emulator's output epilogue.  Replaces, for one spectrum, the gather * window and
 * segment sum of the reference likelihood (tf_planck2018_lite.py:281-291; tables :126-202) at the point where the
 * spectrum leaves the tensor core.  `weight[n]` is the window weight of mode n and `bin_of_mode[n]` its bin (-1: in no
 * bin); bins ascend with n and a group of four consecutive modes touches at most two bins (plik-lite's narrowest bin is
 * five multipoles wide).  A later cpb200_forward_device call with CPB200_FLAG_TEN_TO | CPB200_FLAG_BINNED writes, per
 * row, cpb200_engine_binned_width() floats instead of n_modes: for every group g of four modes the two partial sums
 *   out[row][2g]   = sum of weight[n] * 10**y[n] over the group's modes that fall into the FIRST bin the group touches,
 *   out[row][2g+1] = the same for the SECOND bin
 * (fixed order; half the bytes of the spectrum).  cpb200_plik objects created with `partials` set consume that layout.
 * The output buffer must be 16-byte aligned with a row pitch that is a multiple of 4 floats.
 */
int cpb200_engine_set_binning(cpb200_engine* e, const float* weight, const int32_t* bin_of_mode);
int64_t cpb200_engine_binned_width(const cpb200_engine* e);

/* Pinned host memory for zero-bounce cpb200_forward_host calls. */
int cpb200_host_alloc(void** ptr, uint64_t bytes);
int cpb200_host_free(void* ptr);

/* Introspection used by bench.py / tests. */
int cpb200_engine_info(const cpb200_engine* e, int32_t* n_parameters, int32_t* n_modes,
                       int32_t* precision, int32_t* device);
/* Kernels launched by this engine since creation (each grid launch counts once). */
int64_t cpb200_engine_launch_count(const cpb200_engine* e);
/* Device time of the most recent cpb200_forward_device/_host kernel section, in ms
 * (CUDA events recorded on the launching stream); < 0 if none. Synchronises on those events. */
float cpb200_engine_last_kernel_ms(cpb200_engine* e);
/* SM cycles the slowest CTA of the most recent fused-kernel launch spent between entry and exit (a clock-independent
 * cost figure: wall time also depends on the SM clock the power cap leaves); -1 if none or if timing is off.
 * Synchronises the device. */
int64_t cpb200_engine_last_kernel_cycles(cpb200_engine* e);
/* Turn per-call kernel timing on/off (off by default: two event records, a 8-byte memset and one atomic per CTA
 * per call). */
int cpb200_engine_set_timing(cpb200_engine* e, int enabled);
/*
 * Status words the fused kernel writes into mapped host memory (readable after a device fault):
 *   watchdog_code : 0, or role*100 + barrier of a pipeline role that was stuck for ~3 s before the kernel trapped
 *                   (1xx loader, 2xx MMA issuer / relay, 3xx standardiser, 4xx epilogue); returns CPB200_E_CUDA then.
 *   range_flag    : 1 if, since it was last cleared, a parameter row was beyond +-64 standard deviations of the
 *                   training mean (or not finite) or an activation left the fp16 split range - that batch's result
 *                   is finite but wrong (the operands saturate); cpb200_forward_host checks it itself and returns
 *                   CPB200_E_RANGE, callers of the asynchronous cpb200_forward_device ask here after synchronising.
 * The reference evaluates such rows in float64 without complaint (cosmopower_NN.py:371-384): the Python binding
 * re-runs the batch on the CPB200_PREC_FP32_SIMT kernels when it sees CPB200_E_RANGE.
 */
int cpb200_engine_status(cpb200_engine* e, int32_t* watchdog_code, int32_t* range_flag, int clear_range);

/* Debug: per-CTA cycle counters of the fused kernel's roles (16 int64 per SM; layout UM_CNT_* in
 * csrc/umma_kernel.cuh).  enable!=0 switches collection on for later launches; when `out` is
 * non-NULL the counters of the most recent launch are copied out first (device synchronised).
 * Returns the number of int64 values available, or a negative error. */
int cpb200_engine_debug_counters(cpb200_engine* e, int enable, int64_t* out, int32_t capacity);

/* Debug, needs no device: the fused kernel's job table (the static schedule of one tile pair, see csrc/cpb200_api.cu
 * build_jobs) for a network given by its contractions' logical K, N and epilogue kinds (0 activation, 1 affine, 2 output,
 * 3 quadratic form).  Writes up to `capacity` encoded entries (csrc/umma_kernel.cuh um_job_w0 / um_job_w1) and, in `info`
 * (3 + 3 * n_gemms ints): schedule, rot_mul, parity-buffer stages, then k_stages / chunk width / chunks per contraction.
 * `model` (may be NULL) gives the emulator index of every contraction of a multi-emulator engine.
 * Returns the number of entries or a negative error.  Used by tests/test_job_table.py and tools/schedule_sim.py. */
int cpb200_debug_job_table(int32_t n_gemms, const int32_t* k, const int32_t* n, const int32_t* epi_kind, const int32_t* model,
                           int32_t tri, uint32_t* w0, uint32_t* w1, int32_t capacity, int32_t* info);

/* -------------------------------------------------------------------------------------------------
 * Planck 2018 plik-lite likelihood tail behind the TT / TE / EE emulators.
 *
 * Replaces: tf_planck2018_lite.get_loglkl from the point where the three spectra exist
 * (cosmopower/likelihoods/tf_planck2018_lite/tf_planck2018_lite.py:281-304): units factor,
 * gather * window and segment sum into the band powers (:284-291; tables built at :126-202),
 * division by A_planck^2 (:294), difference to the data vector and Fisher quadratic form (:297-302),
 * -chi^2 / 2 (:302).  The Fisher matrix (:206-230) is computed by the caller and passed in.
 * ------------------------------------------------------------------------------------------------- */
typedef struct cpb200_plik cpb200_plik;

typedef struct {
    int32_t n_ell;              /* multipoles per spectrum row (2507: ell = 2 .. 2508) */
    int32_t n_bins;             /* 613 = 215 TT + 199 TE + 199 EE */
    int32_t n_gather;           /* total gathered multipoles (tf_planck2018_lite.py:196-197), = bin_start[n_bins] */
    const int32_t* bin_start;   /* n_bins + 1 offsets into `window` (CSR over the reference's indices_rep) */
    const int32_t* gather_first;/* n_bins: position of a bin's first multipole in the concatenated [TT|TE|EE] row;
                                   a bin gathers bin_start[b+1]-bin_start[b] consecutive multipoles (:131-176) */
    const float* window;        /* n_gather window weights, the reference's window_ttteee (:282) */
    const float* x_data;        /* n_bins data vector (:107-109) */
    const float* fisher;        /* n_bins x n_bins, row-major, symmetric (:206-230) */
    float units_factor;         /* (T_CMB in muK)^2 (:45, :281) */
    int32_t te_binned;          /* != 0: the TE input of cpb200_plik_loglike_device holds that spectrum's band powers (bins
                                   215 .. 413 of the vector, before units and calibration) instead of multipoles - for a TE
                                   emulator whose linear PCA basis was folded with the window functions on the host */
    int32_t partials;           /* != 0: the TT and EE inputs hold the partial band powers written by emulators with
                                   cpb200_engine_set_binning + CPB200_FLAG_BINNED (2 * ceil(n_ell / 4) floats per row, pitch
                                   `pitch` of cpb200_plik_loglike_device) instead of multipoles: the binning pass then only
                                   adds a bin's partials */
} cpb200_plik_desc;

int cpb200_plik_create(const cpb200_plik_desc* desc, int32_t device, cpb200_plik** out);
int cpb200_plik_destroy(cpb200_plik* p);

/*
 * loglike_dev[r] = -0.5 * chi^2 of row r, for r < n_rows.  cl_tt / cl_te / cl_ee are device float32
 * (n_rows, n_ell) arrays with a common row pitch in elements (the emulators' outputs: TT and EE from
 * ten_to_predictions, TE from predictions, :276-278; TE has its own row pitch te_pitch, and holds band powers when
 * the object was created with te_binned); cal_dev holds A_planck per row (:270).
 * binned_dev, when non-NULL, receives the (n_rows, n_bins) model band powers X_model (:294) with row pitch
 * binned_pitch.  Asynchronous on `stream`; rows are processed in engine-owned chunks.
 */
int cpb200_plik_loglike_device(cpb200_plik* p, const float* cl_tt_dev, const float* cl_te_dev,
                               const float* cl_ee_dev, int64_t pitch, int64_t te_pitch, const float* cal_dev,
                               int64_t n_rows, float* loglike_dev, float* binned_dev, int64_t binned_pitch,
                               void* stream);
/* Kernels launched by this object since creation. */
int64_t cpb200_plik_launch_count(const cpb200_plik* p);

/* Message for the most recent failure on this thread ("" if none). */
const char* cpb200_last_error(void);

#ifdef __cplusplus
}
#endif
#endif /* CPB200_H */
