"""ctypes binding of the C ABI declared in include/cpb200.h (csrc/libcpb200.so).

There is deliberately no fallback: if the shared library is missing or no sm_100 device is
visible, creating an engine raises.  Nothing here computes on the CPU.
"""
import ctypes
import os
import threading

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("CPB200_LIB") or os.path.join(_HERE, "csrc", "libcpb200.so")

PREC_FP32_SIMT = 0
PREC_F16X3 = 1
FLAG_TEN_TO = 1
FLAG_ACCUMULATE = 2
FLAG_BINNED = 4
PRECISIONS = {"fp32": PREC_FP32_SIMT, "fp32_simt": PREC_FP32_SIMT, "f16x3": PREC_F16X3}

_fp = ctypes.POINTER(ctypes.c_float)
_fpp = ctypes.POINTER(_fp)


class ModelDesc(ctypes.Structure):
    _fields_ = [("n_parameters", ctypes.c_int32), ("n_modes", ctypes.c_int32),
                ("n_dense", ctypes.c_int32), ("n_pcas", ctypes.c_int32),
                ("dims", ctypes.POINTER(ctypes.c_int32)),
                ("W", _fpp), ("b", _fpp), ("alphas", _fpp), ("betas", _fpp),
                ("parameters_mean", _fp), ("parameters_std", _fp),
                ("features_mean", _fp), ("features_std", _fp),
                ("pca_mean", _fp), ("pca_std", _fp), ("pca_transform_matrix", _fp)]


class PlikDesc(ctypes.Structure):
    _fields_ = [("n_ell", ctypes.c_int32), ("n_bins", ctypes.c_int32), ("n_gather", ctypes.c_int32),
                ("bin_start", ctypes.POINTER(ctypes.c_int32)), ("gather_first", ctypes.POINTER(ctypes.c_int32)),
                ("window", _fp), ("x_data", _fp), ("fisher", _fp), ("units_factor", ctypes.c_float), ("te_binned", ctypes.c_int32),
                ("partials", ctypes.c_int32)]


# every symbol include/cpb200.h declares: (name, restype, argtypes)
SYMBOLS = [
    ("cpb200_abi_version", ctypes.c_int, []),
    ("cpb200_device_count", ctypes.c_int, []),
    ("cpb200_create", ctypes.c_int, [ctypes.POINTER(ModelDesc), ctypes.c_int, ctypes.c_int,
                                     ctypes.POINTER(ctypes.c_void_p)]),
    ("cpb200_destroy", ctypes.c_int, [ctypes.c_void_p]),
    ("cpb200_forward_device", ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64,
                                             ctypes.c_void_p, ctypes.c_int64, ctypes.c_int, ctypes.c_void_p]),
    ("cpb200_forward_host", ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64,
                                           ctypes.c_void_p, ctypes.c_int64, ctypes.c_int]),
    ("cpb200_forward_host_f64", ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p, ctypes.c_int64, ctypes.c_int]),
    ("cpb200_forward_host_multi", ctypes.c_int, [ctypes.POINTER(ctypes.c_void_p), ctypes.c_int32, ctypes.c_void_p, ctypes.c_int64,
                                                 ctypes.c_void_p, ctypes.c_int64, ctypes.c_int]),
    ("cpb200_forward_host_multi_f64", ctypes.c_int, [ctypes.POINTER(ctypes.c_void_p), ctypes.c_int32, ctypes.c_void_p, ctypes.c_int64,
                                                     ctypes.c_void_p, ctypes.c_int64, ctypes.c_int]),
    ("cpb200_host_alloc", ctypes.c_int, [ctypes.POINTER(ctypes.c_void_p), ctypes.c_uint64]),
    ("cpb200_host_free", ctypes.c_int, [ctypes.c_void_p]),
    ("cpb200_engine_info", ctypes.c_int, [ctypes.c_void_p] + [ctypes.POINTER(ctypes.c_int32)] * 4),
    ("cpb200_engine_launch_count", ctypes.c_int64, [ctypes.c_void_p]),
    ("cpb200_engine_last_kernel_ms", ctypes.c_float, [ctypes.c_void_p]),
    ("cpb200_engine_last_kernel_cycles", ctypes.c_int64, [ctypes.c_void_p]),
    ("cpb200_engine_set_timing", ctypes.c_int, [ctypes.c_void_p, ctypes.c_int]),
    ("cpb200_engine_set_binning", ctypes.c_int, [ctypes.c_void_p, _fp, ctypes.POINTER(ctypes.c_int32)]),
    ("cpb200_engine_binned_width", ctypes.c_int64, [ctypes.c_void_p]),
    ("cpb200_engine_status", ctypes.c_int, [ctypes.c_void_p, ctypes.POINTER(ctypes.c_int32), ctypes.POINTER(ctypes.c_int32), ctypes.c_int]),
    ("cpb200_engine_debug_counters", ctypes.c_int, [ctypes.c_void_p, ctypes.c_int,
                                                    ctypes.POINTER(ctypes.c_int64), ctypes.c_int32]),
    ("cpb200_debug_job_table", ctypes.c_int, [ctypes.c_int32, ctypes.POINTER(ctypes.c_int32), ctypes.POINTER(ctypes.c_int32),
                                              ctypes.POINTER(ctypes.c_int32), ctypes.POINTER(ctypes.c_int32), ctypes.c_int32,
                                              ctypes.POINTER(ctypes.c_uint32), ctypes.POINTER(ctypes.c_uint32), ctypes.c_int32,
                                              ctypes.POINTER(ctypes.c_int32)]),
    ("cpb200_create_multi", ctypes.c_int, [ctypes.POINTER(ctypes.POINTER(ModelDesc)), ctypes.c_int32, ctypes.c_int,
                                           ctypes.POINTER(ctypes.c_void_p)]),
    ("cpb200_forward_multi_device", ctypes.c_int, [ctypes.c_void_p, ctypes.POINTER(ctypes.c_void_p), ctypes.c_int64,
                                                   ctypes.POINTER(ctypes.c_void_p), ctypes.POINTER(ctypes.c_int64),
                                                   ctypes.POINTER(ctypes.c_int32), ctypes.c_void_p]),
    ("cpb200_plik_create", ctypes.c_int, [ctypes.POINTER(PlikDesc), ctypes.c_int32, ctypes.POINTER(ctypes.c_void_p)]),
    ("cpb200_plik_destroy", ctypes.c_int, [ctypes.c_void_p]),
    ("cpb200_plik_loglike_device", ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                                  ctypes.c_int64, ctypes.c_int64, ctypes.c_void_p, ctypes.c_int64,
                                                  ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p]),
    ("cpb200_plik_launch_count", ctypes.c_int64, [ctypes.c_void_p]),
    ("cpb200_last_error", ctypes.c_char_p, []),
]

_lib = None
_lock = threading.Lock()


def load_library():
    """Load csrc/libcpb200.so and bind every declared symbol; raises if it is not built."""
    global _lib
    with _lock:
        if _lib is None:
            if not os.path.isfile(LIB_PATH):
                raise RuntimeError(
                    "cosmopower_b200: %s is missing - build it with `python -c 'import __graft_entry__ as g; "
                    "g.build()'` or `make -C cosmopower_b200/csrc`. There is no CPU fallback." % LIB_PATH)
            lib = ctypes.CDLL(LIB_PATH)
            for name, res, args in SYMBOLS:
                fn = getattr(lib, name)
                fn.restype = res
                fn.argtypes = args
            _lib = lib
    return _lib


E_RANGE = -5


class Cpb200Error(RuntimeError):
    """A non-zero return code of the C ABI; `.code` is the CPB200_E_* value."""

    def __init__(self, code, msg):
        super().__init__("cpb200 error %d: %s" % (code, msg))
        self.code = code


def _check(rc):
    if rc != 0:
        msg = load_library().cpb200_last_error().decode("utf-8", "replace")
        raise Cpb200Error(rc, msg)


def _f32(a):
    return np.ascontiguousarray(np.asarray(a, dtype=np.float32))


def _ptr(a):
    return a.ctypes.data_as(_fp)


def job_table(dims_k, dims_n, kinds, tri=False, models=None):
    """The fused kernel's static schedule for a network (cpb200_debug_job_table; no device needed).
    Returns (entries, info): entries = list of dicts (g, t, c, nxt, s0, ns, first, last, slot, two, in_code, out_code),
    info = dict(schedule, rot_mul, par_stages, gemms=[(k_stages, nc, n_chunks)])."""
    lib = load_library()
    n = len(dims_k)
    arr = lambda v: (ctypes.c_int32 * len(v))(*[int(x) for x in v])
    cap = 512
    w0, w1 = (ctypes.c_uint32 * cap)(), (ctypes.c_uint32 * cap)()
    info = (ctypes.c_int32 * (3 + 3 * n))()
    cnt = lib.cpb200_debug_job_table(n, arr(dims_k), arr(dims_n), arr(kinds), arr(models) if models is not None else None,
                                     1 if tri else 0, w0, w1, cap, info)
    if cnt < 0:
        _check(cnt)
    entries = []
    for i in range(cnt):
        a, b = int(w0[i]), int(w1[i])
        entries.append(dict(g=a & 15, t=(a >> 4) & 1, c=(a >> 5) & 127, nxt=(a >> 12) & 1, s0=(a >> 13) & 63, ns=(a >> 19) & 63,
                            first=(a >> 25) & 1, last=(a >> 26) & 1, slot=(a >> 27) & 3, two=(a >> 29) & 1,
                            in_code=b & 0xff, out_code=(b >> 8) & 0xff))
    return entries, dict(schedule=int(info[0]), rot_mul=int(info[1]), par_stages=int(info[2]),
                         gemms=[(int(info[3 + 3 * i]), int(info[4 + 3 * i]), int(info[5 + 3 * i])) for i in range(n)])


class PinnedArray:
    """A float32 host array in page-locked memory (cpb200_host_alloc); `.array` is the numpy view."""

    def __init__(self, shape):
        lib = load_library()
        self._shape = tuple(int(s) for s in shape)
        n = int(np.prod(self._shape)) if self._shape else 1
        p = ctypes.c_void_p()
        _check(lib.cpb200_host_alloc(ctypes.byref(p), ctypes.c_uint64(max(n, 1) * 4)))
        self._ptr = p
        buf = (ctypes.c_float * max(n, 1)).from_address(p.value)
        self.array = np.frombuffer(buf, dtype=np.float32, count=n).reshape(self._shape)

    def free(self):
        if self._ptr is not None:
            self.array = None
            load_library().cpb200_host_free(self._ptr)
            self._ptr = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


def _model_desc(attrs, keep):
    """cpb200_model_desc of one restored emulator; the arrays it points to are appended to `keep`."""
    n_dense = len(attrs["W_"])
    dims = [int(attrs["W_"][0].shape[0])] + [int(w.shape[1]) for w in attrs["W_"]]

    def arr_list(items):
        arrs = [_f32(a) for a in items]
        keep.extend(arrs)
        c = (_fp * max(len(arrs), 1))(*[_ptr(a) for a in arrs])
        keep.append(c)
        return c

    def arr(a):
        a = _f32(a)
        keep.append(a)
        return _ptr(a)

    d = ModelDesc()
    d.n_parameters = int(attrs["n_parameters"])
    d.n_modes = int(attrs["n_modes"])
    d.n_dense = n_dense
    d.n_pcas = int(attrs.get("n_pcas", 0) or 0)
    cdims = (ctypes.c_int32 * len(dims))(*dims)
    keep.append(cdims)
    d.dims = cdims
    for i, w in enumerate(attrs["W_"]):
        if tuple(w.shape) != (dims[i], dims[i + 1]):
            raise ValueError("W_[%d] has shape %s, expected %s" % (i, w.shape, (dims[i], dims[i + 1])))
    d.W = arr_list(attrs["W_"])
    d.b = arr_list(attrs["b_"])
    d.alphas = arr_list(attrs["alphas_"])
    d.betas = arr_list(attrs["betas_"])
    d.parameters_mean = arr(attrs["parameters_mean_"])
    d.parameters_std = arr(attrs["parameters_std_"])
    d.features_mean = arr(attrs["features_mean_"])
    d.features_std = arr(attrs["features_std_"])
    if d.n_pcas:
        d.pca_mean = arr(attrs["pca_mean_"])
        d.pca_std = arr(attrs["pca_std_"])
        d.pca_transform_matrix = arr(attrs["pca_transform_matrix_"])
    return d


class Engine:
    """One restored emulator resident on one B200 (opaque cpb200_engine handle)."""

    def __init__(self, attrs, device=0, precision="f16x3"):
        lib = load_library()
        self._lib = lib
        self._h = None
        keep = []
        d = _model_desc(attrs, keep)
        prec = PRECISIONS[precision] if isinstance(precision, str) else int(precision)
        h = ctypes.c_void_p()
        _check(lib.cpb200_create(ctypes.byref(d), int(device), prec, ctypes.byref(h)))
        self._h = h
        self.device = int(device)
        self.precision = prec
        self.n_parameters = d.n_parameters
        self.n_modes = d.n_modes
        # the C ABI asks callers to serialise the host threads that use one engine (launches themselves are
        # ordered across streams by the library): every call below holds this lock
        self._call_lock = threading.RLock()

    def close(self):
        if self._h is not None:
            with self._call_lock:
                self._lib.cpb200_destroy(self._h)
                self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- device-resident path -------------------------------------------------------------
    def set_binning(self, weight, bin_of_mode):
        """Attach band-power binning to the output epilogue (cpb200_engine_set_binning): weight[n] = window weight of mode
        n, bin_of_mode[n] = its bin or -1.  forward_torch(..., ten_to=True, binned=True) then returns the partial band
        powers (N, binned_width) instead of the spectra."""
        w = _f32(weight)
        b = np.ascontiguousarray(bin_of_mode, dtype=np.int32)
        if w.shape != (self.n_modes,) or b.shape != (self.n_modes,):
            raise ValueError("weight and bin_of_mode must have n_modes entries")
        with self._call_lock:
            _check(self._lib.cpb200_engine_set_binning(self._h, _ptr(w), b.ctypes.data_as(ctypes.POINTER(ctypes.c_int32))))
        self.binned_width = int(self._lib.cpb200_engine_binned_width(self._h))

    def forward_device_ptr(self, params_ptr, n_rows, out_ptr, out_pitch, ten_to=False, stream=0, accumulate=False, binned=False):
        flags = (FLAG_TEN_TO if ten_to else 0) | (FLAG_ACCUMULATE if accumulate else 0) | (FLAG_BINNED if binned else 0)
        with self._call_lock:
            _check(self._lib.cpb200_forward_device(self._h, ctypes.c_void_p(params_ptr), int(n_rows),
                                                   ctypes.c_void_p(out_ptr), int(out_pitch), flags, ctypes.c_void_p(stream)))

    def forward_torch(self, params, out=None, ten_to=False, accumulate=False, binned=False):
        """params: CUDA float32 tensor (N, P) on this engine's device -> CUDA float32 (N, n_modes); with binned=True (after
        set_binning, with ten_to=True) -> (N, binned_width) partial band powers of 10**y."""
        import torch
        if binned:
            if not (params.is_cuda and params.dtype == torch.float32 and params.dim() == 2 and params.shape[1] == self.n_parameters
                    and params.device.index == self.device):
                raise ValueError("expected a CUDA float32 tensor of shape (N, %d) on device %d" % (self.n_parameters, self.device))
            params = params.contiguous()
            n = params.shape[0]
            out = torch.empty((n, self.binned_width), dtype=torch.float32, device=params.device)
            stream = torch.cuda.current_stream(params.device).cuda_stream
            self.forward_device_ptr(params.data_ptr(), n, out.data_ptr(), self.binned_width, ten_to, stream, False, True)
            return out
        if not (params.is_cuda and params.dtype == torch.float32 and params.dim() == 2
                and params.shape[1] == self.n_parameters):
            raise ValueError("expected a CUDA float32 tensor of shape (N, %d)" % self.n_parameters)
        if params.device.index != self.device:
            raise ValueError("tensor is on cuda:%s, engine on cuda:%d" % (params.device.index, self.device))
        params = params.contiguous()
        n = params.shape[0]
        if accumulate and out is None:
            raise ValueError("accumulate=True needs the `out` tensor that holds the first emulator's log-spectra")
        if out is None:
            # rows padded to a multiple of 8 floats (32-byte sectors) so the kernel can use aligned
            # vector stores; the caller gets the (n, n_modes) view
            pitch = (self.n_modes + 7) // 8 * 8
            out = torch.empty((n, pitch), dtype=torch.float32, device=params.device)[:, :self.n_modes]
        elif not (out.is_cuda and out.dtype == torch.float32 and out.dim() == 2 and out.shape[0] == n
                  and out.shape[1] == self.n_modes and out.stride(1) == 1 and out.stride(0) >= self.n_modes):
            raise ValueError("bad output tensor")
        stream = torch.cuda.current_stream(params.device).cuda_stream
        self.forward_device_ptr(params.data_ptr(), n, out.data_ptr(), out.stride(0) if n > 1 else self.n_modes,
                                ten_to, stream, accumulate)
        return out

    # -- host path --------------------------------------------------------------------------
    def forward_host(self, params, out=None, ten_to=False, dtype=np.float32):
        """params: float32 C-contiguous (N, P) host array -> (N, n_modes) host array, float32 or - widened by the
        library's host threads on the way out - float64 (`dtype`, or the dtype of `out`)."""
        params = np.ascontiguousarray(params, dtype=np.float32)
        if params.ndim != 2 or params.shape[1] != self.n_parameters:
            raise ValueError("expected an array of shape (N, %d), got %s" % (self.n_parameters, params.shape))
        n = params.shape[0]
        if out is None:
            out = np.empty((n, self.n_modes), dtype=dtype)
        if out.dtype not in (np.float32, np.float64) or out.ndim != 2 or out.shape != (n, self.n_modes) or (n > 0 and out.strides[1] != out.itemsize):
            raise ValueError("out must be a float32 or float64 (N, %d) array with contiguous rows" % self.n_modes)
        fn = self._lib.cpb200_forward_host if out.dtype == np.float32 else self._lib.cpb200_forward_host_f64
        with self._call_lock:
            _check(fn(self._h, ctypes.c_void_p(params.ctypes.data), n, ctypes.c_void_p(out.ctypes.data),
                      out.strides[0] // out.itemsize if n > 1 else self.n_modes, FLAG_TEN_TO if ten_to else 0))
        return out

    def status(self, clear_range=False):
        """(watchdog_code, range_flag) of the fused kernel (cpb200_engine_status); after an asynchronous
        forward_torch, call it once the stream is synchronised: range_flag == 1 means a row was outside the
        fp16 split range and that batch must be recomputed with precision="fp32"."""
        w, r = ctypes.c_int32(0), ctypes.c_int32(0)
        with self._call_lock:
            _check(self._lib.cpb200_engine_status(self._h, ctypes.byref(w), ctypes.byref(r), 1 if clear_range else 0))
        return int(w.value), int(r.value)

    # -- introspection ---------------------------------------------------------------------------
    def launch_count(self):
        return int(self._lib.cpb200_engine_launch_count(self._h))

    def set_timing(self, enabled):
        _check(self._lib.cpb200_engine_set_timing(self._h, 1 if enabled else 0))

    def debug_counters(self, enable=True, read=False):
        """Per-CTA role cycle counters of the fused kernel (debug aid; see csrc/umma_kernel.cuh UM_CNT_*)."""
        if not read:
            n = self._lib.cpb200_engine_debug_counters(self._h, 1 if enable else 0, None, 0)
            if n < 0:
                _check(n)
            return None
        n = self._lib.cpb200_engine_debug_counters(self._h, 1, None, 0)
        buf = (ctypes.c_int64 * n)()
        n = self._lib.cpb200_engine_debug_counters(self._h, 1 if enable else 0, buf, n)
        if n < 0:
            _check(n)
        return np.frombuffer(buf, dtype=np.int64).reshape(-1, 16).copy()

    def last_kernel_cycles(self):
        return int(self._lib.cpb200_engine_last_kernel_cycles(self._h))

    def last_kernel_ms(self):
        return float(self._lib.cpb200_engine_last_kernel_ms(self._h))


class MultiEngine:
    """Several emulators evaluated on the same rows by ONE launch (cpb200_create_multi / cpb200_forward_multi_device):
    TT + TE + EE behind the likelihood, or the linear P(k) and its nonlinear boost with the log-sum fused."""

    def __init__(self, attrs_list, device=0):
        lib = load_library()
        self._lib = lib
        self._h = None
        keep = []
        descs = [_model_desc(a, keep) for a in attrs_list]
        ptrs = (ctypes.POINTER(ModelDesc) * len(descs))(*[ctypes.pointer(d) for d in descs])
        h = ctypes.c_void_p()
        _check(lib.cpb200_create_multi(ptrs, len(descs), int(device), ctypes.byref(h)))
        self._h = h
        self.device = int(device)
        self.n_models = len(descs)
        self.n_parameters = [int(d.n_parameters) for d in descs]
        self.n_modes = [int(d.n_modes) for d in descs]
        self._call_lock = threading.RLock()

    def close(self):
        if self._h is not None:
            with self._call_lock:
                self._lib.cpb200_destroy(self._h)
                self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def forward_torch(self, params, outs=None, ten_to=False, accumulate=False):
        """params: one CUDA float32 (N, P_m) tensor per emulator (the same tensor may be passed several times);
        ten_to / accumulate: one bool per emulator (or one for all).  With accumulate[m] emulator m adds its log10
        spectrum to emulator m-1's and writes into the same tensor.  Returns the list of output tensors."""
        import torch
        M = self.n_models
        if len(params) != M:
            raise ValueError("need one parameter tensor per emulator")
        ten_to = [bool(ten_to)] * M if isinstance(ten_to, (bool, int)) else [bool(v) for v in ten_to]
        accumulate = [bool(accumulate and m > 0) for m in range(M)] if isinstance(accumulate, (bool, int)) else [bool(v) for v in accumulate]
        n = int(params[0].shape[0])
        xs = []
        for m, x in enumerate(params):
            if not (x.is_cuda and x.dtype == torch.float32 and x.dim() == 2 and x.shape == (n, self.n_parameters[m]) and x.device.index == self.device):
                raise ValueError("emulator %d expects a CUDA float32 tensor of shape (%d, %d) on device %d" % (m, n, self.n_parameters[m], self.device))
            xs.append(x.contiguous())
        if outs is None:
            outs = []
            for m in range(M):
                if accumulate[m]:
                    outs.append(outs[m - 1])
                else:
                    pitch = (self.n_modes[m] + 7) // 8 * 8
                    outs.append(torch.empty((n, pitch), dtype=torch.float32, device=xs[0].device)[:, :self.n_modes[m]])
        for m, o in enumerate(outs):
            if not (o.is_cuda and o.dtype == torch.float32 and o.dim() == 2 and o.shape == (n, self.n_modes[m]) and o.stride(1) == 1):
                raise ValueError("bad output tensor for emulator %d" % m)
        pp = (ctypes.c_void_p * M)(*[x.data_ptr() for x in xs])
        po = (ctypes.c_void_p * M)(*[o.data_ptr() for o in outs])
        pitch = (ctypes.c_int64 * M)(*[int(o.stride(0)) if n > 1 else self.n_modes[m] for m, o in enumerate(outs)])
        fl = (ctypes.c_int32 * M)(*[(FLAG_TEN_TO if ten_to[m] else 0) | (FLAG_ACCUMULATE if accumulate[m] else 0) for m in range(M)])
        stream = torch.cuda.current_stream(xs[0].device).cuda_stream
        with self._call_lock:
            _check(self._lib.cpb200_forward_multi_device(self._h, pp, n, po, pitch, fl, ctypes.c_void_p(stream)))
        return outs

    launch_count = Engine.launch_count
    set_timing = Engine.set_timing
    debug_counters = Engine.debug_counters
    last_kernel_cycles = Engine.last_kernel_cycles
    last_kernel_ms = Engine.last_kernel_ms


def forward_host_multi(engines, params, out=None, ten_to=False, dtype=np.float32):
    """One batch over several devices from one process (cpb200_forward_host_multi): `engines` holds one Engine of the
    same model per device; rows are sharded contiguously, one library thread per device.  Same result as
    engines[0].forward_host(params), bit for bit."""
    if len(engines) == 1:
        return engines[0].forward_host(params, out=out, ten_to=ten_to, dtype=dtype)
    e0 = engines[0]
    params = np.ascontiguousarray(params, dtype=np.float32)
    if params.ndim != 2 or params.shape[1] != e0.n_parameters:
        raise ValueError("expected an array of shape (N, %d), got %s" % (e0.n_parameters, params.shape))
    n = params.shape[0]
    if out is None:
        out = np.empty((n, e0.n_modes), dtype=dtype)
    if out.dtype not in (np.float32, np.float64) or out.ndim != 2 or out.shape != (n, e0.n_modes) or (n > 0 and out.strides[1] != out.itemsize):
        raise ValueError("out must be a float32 or float64 (N, %d) array with contiguous rows" % e0.n_modes)
    lib = e0._lib
    fn = lib.cpb200_forward_host_multi if out.dtype == np.float32 else lib.cpb200_forward_host_multi_f64
    handles = (ctypes.c_void_p * len(engines))(*[e._h for e in engines])
    locks = sorted(engines, key=id)
    for e in locks:
        e._call_lock.acquire()
    try:
        _check(fn(handles, len(engines), ctypes.c_void_p(params.ctypes.data), n, ctypes.c_void_p(out.ctypes.data),
                  out.strides[0] // out.itemsize if n > 1 else e0.n_modes, FLAG_TEN_TO if ten_to else 0))
    finally:
        for e in reversed(locks):
            e._call_lock.release()
    return out


class PlikEngine:
    """Device-side Planck-lite likelihood tail (cpb200_plik_*): band-power binning + Fisher quadratic form."""

    def __init__(self, n_ell, bin_start, gather_first, window, x_data, fisher, units_factor, device=0, te_binned=False, partials=False):
        lib = load_library()
        self._lib = lib
        self._keep = [np.ascontiguousarray(bin_start, dtype=np.int32), np.ascontiguousarray(gather_first, dtype=np.int32),
                      _f32(window), _f32(x_data), _f32(fisher)]
        bs, gf, w, xd, f = self._keep
        self.n_ell, self.n_bins, self.device = int(n_ell), int(xd.size), int(device)
        if f.shape != (self.n_bins, self.n_bins) or bs.size != self.n_bins + 1 or gf.size != self.n_bins:
            raise ValueError("inconsistent likelihood tables")
        ip = ctypes.POINTER(ctypes.c_int32)
        desc = PlikDesc(self.n_ell, self.n_bins, int(w.size), bs.ctypes.data_as(ip), gf.ctypes.data_as(ip),
                        _ptr(w), _ptr(xd), _ptr(f), float(units_factor), 1 if te_binned else 0, 1 if partials else 0)
        self.te_binned = bool(te_binned)
        self.partials = bool(partials)
        self.tt_ee_width = 2 * ((self.n_ell + 3) // 4) if partials else self.n_ell
        self.n_te_bins = int(np.sum(gf // self.n_ell == 1))
        h = ctypes.c_void_p()
        _check(lib.cpb200_plik_create(ctypes.byref(desc), self.device, ctypes.byref(h)))
        self._h = h

    def close(self):
        if getattr(self, "_h", None):
            self._lib.cpb200_plik_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def launch_count(self):
        return int(self._lib.cpb200_plik_launch_count(self._h))

    def loglike_torch(self, cl_tt, cl_te, cl_ee, cal, return_binned=False):
        """cl_* : CUDA float32 (N, n_ell) tensors with a common row stride; cal: CUDA float32 (N,).
        Returns loglike (N,) [and the (N, n_bins) model band powers]."""
        import torch
        n = int(cl_tt.shape[0])
        for t, width in ((cl_tt, self.tt_ee_width), (cl_te, self.n_te_bins if self.te_binned else self.n_ell), (cl_ee, self.tt_ee_width)):
            if not (t.is_cuda and t.dtype == torch.float32 and t.dim() == 2 and t.shape[0] == n and t.shape[1] >= width and
                    t.stride(1) == 1 and t.device.index == self.device):
                raise ValueError("inputs must be CUDA float32 (N, %d) tensors on device %d" % (width, self.device))
        if cl_tt.stride(0) != cl_ee.stride(0):
            raise ValueError("the TT and EE inputs must share one row stride")
        cal = cal.to(device=cl_tt.device, dtype=torch.float32).contiguous().reshape(-1)
        if cal.numel() != n:
            raise ValueError("cal must hold one A_planck per row")
        out = torch.empty((n,), dtype=torch.float32, device=cl_tt.device)
        binned = torch.empty((n, self.n_bins), dtype=torch.float32, device=cl_tt.device) if return_binned else None
        stream = torch.cuda.current_stream(cl_tt.device).cuda_stream
        _check(self._lib.cpb200_plik_loglike_device(
            self._h, cl_tt.data_ptr(), cl_te.data_ptr(), cl_ee.data_ptr(), int(cl_tt.stride(0)) if n else self.n_ell,
            int(cl_te.stride(0)) if n else self.n_ell, cal.data_ptr(), n, out.data_ptr(), binned.data_ptr() if return_binned else None,
            self.n_bins, ctypes.c_void_p(stream)))
        return (out, binned) if return_binned else out
