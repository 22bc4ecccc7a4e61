"""Shared host-side behaviour of the two drop-in classes (restore, dict ordering, dtype rules)."""
import os

import numpy as np

from . import _engine, _loader


class _EmulatorBase:
    _FIELDS = ()
    _IS_PCA = False

    # -- construction -------------------------------------------------------------------------
    def _init_runtime(self, device, precision, verbose):
        self._device = device
        self._precision = precision
        self._engines = {}
        self._engines_fp32 = {}
        self.verbose = verbose

    def restore(self, filename):
        """Load a pre-trained model from `filename + ".pkl"` (reference layout; TF-free reader),
        or from `filename + ".npz"` (flat container) when no pickle exists."""
        if os.path.isfile(filename + ".pkl") or not os.path.isfile(filename + ".npz"):
            # open() raises FileNotFoundError exactly like the reference when the pickle is missing
            items = _loader.load_pickle_list(filename + ".pkl")
            attrs = _loader.attributes_from_list(items, self._FIELDS)
        else:
            attrs = _loader.load_npz(filename + ".npz", self._FIELDS)
        for k, v in attrs.items():
            setattr(self, k, v)
        self._validate()
        self._engines = {}
        self._engines_fp32 = {}

    def _validate(self):
        arch = [int(self.W_[0].shape[0])] + [int(w.shape[1]) for w in self.W_]
        if list(self.architecture) != arch:
            raise ValueError("architecture %s does not match the weight shapes %s" % (list(self.architecture), arch))
        if int(self.n_layers) != len(self.W_):
            raise ValueError("n_layers %s != number of weight matrices %d" % (self.n_layers, len(self.W_)))

    def _attrs(self):
        return {k: getattr(self, k) for k in self._FIELDS}

    def engine(self, device=None):
        """The device-resident engine for this model (created on first use; one per device)."""
        dev = self._device if device is None else device
        if dev is None:
            dev = 0
            try:
                import torch
                if torch.cuda.is_available():
                    dev = torch.cuda.current_device()
            except Exception:
                pass
        if dev not in self._engines:
            if self._precision not in ("f16x3", "fp32", "auto"):
                raise ValueError("precision must be \"f16x3\", \"fp32\" or \"auto\", got %r" % (self._precision,))
            self._engines[dev] = _engine.Engine(self._attrs(), device=dev, precision="f16x3" if self._precision == "auto" else self._precision)
        return self._engines[dev]

    # precision="auto" (extension): batches of up to SMALL_ROWS rows - a sampler's single point - run on the fp32 CUDA-core
    # kernels, whose column-per-warp form takes 35-50 us per call against 100-120 us for the tensor path (one tile walks the
    # whole network whatever the batch); larger batches use the tensor path.  Both are inside the tolerance; they differ in
    # the last bits, which is why the default ("f16x3") keeps every batch size on the same arithmetic.
    SMALL_ROWS = 8

    def _fp32_engine(self, dev):
        if dev not in self._engines_fp32:
            self._engines_fp32[dev] = _engine.Engine(self._attrs(), device=dev, precision="fp32")
        return self._engines_fp32[dev]

    def _engine_for(self, n_rows, device=None):
        eng = self.engine(device)
        if self._precision == "auto" and n_rows <= self.SMALL_ROWS:
            return self._fp32_engine(eng.device)
        return eng

    def close(self):
        for e in list(self._engines.values()) + list(self._engines_fp32.values()):
            e.close()
        self._engines = {}
        self._engines_fp32 = {}

    # -- reference API: NumPy in, NumPy out -----------------------------------------------------------
    def dict_to_ordered_arr_np(self, input_dict):
        """Sort input parameters (reference cosmopower_NN.py:333-350 / cosmopower_PCAplusNN.py:330-347)."""
        if self.parameters is not None:
            return np.stack([input_dict[k] for k in self.parameters], axis=1)
        return np.stack([input_dict[k] for k in input_dict], axis=1)

    def forward_pass_np(self, parameters_arr, devices=None):
        """Forward pass on the B200 (reference cosmopower_NN.py:354-384 / cosmopower_PCAplusNN.py:351-381).
        The contraction runs in split-fp16/fp32 on the device; the returned dtype follows the
        reference's promotion rule (float32 inputs -> float32, anything else -> float64).
        `devices` (extension): a list of CUDA device indices, or "all" - the rows are sharded over those GPUs from this
        one process (one engine and one host thread per device, no collective); the result is the same array."""
        return self._forward_np(parameters_arr, ten_to=False, devices=devices)

    def _device_list(self, devices):
        if devices is None:
            return None
        if isinstance(devices, str):
            if devices != "all":
                raise ValueError("devices must be a list of device indices or \"all\"")
            devices = list(range(_engine.load_library().cpb200_device_count()))
        devices = [int(d) for d in devices]
        if len(set(devices)) != len(devices) or not devices:
            raise ValueError("devices must name each GPU once")
        return devices

    def _forward_np(self, parameters_arr, ten_to, devices=None):
        x = np.asarray(parameters_arr)
        if x.ndim != 2 or x.shape[1] != int(self.n_parameters):
            raise ValueError("parameters array must have shape (N, %d), got %s" % (self.n_parameters, x.shape))
        out_dtype = np.float32 if x.dtype == np.float32 else np.float64
        x32 = np.ascontiguousarray(x, dtype=np.float32)
        # float64 results are widened inside the library while they are copied out (cpb200_forward_host_f64)
        devices = self._device_list(devices)
        try:
            if devices is not None and len(devices) > 1:
                return _engine.forward_host_multi([self.engine(d) for d in devices], x32, ten_to=ten_to, dtype=out_dtype)
            return self._engine_for(x32.shape[0], devices[0] if devices else None).forward_host(x32, ten_to=ten_to, dtype=out_dtype)
        except _engine.Cpb200Error as err:
            if err.code != _engine.E_RANGE:
                raise
        # a row beyond the fp16 split range (|z| >= 64 sigma, a huge activation, NaN): the reference evaluates those
        # without complaint (cosmopower_NN.py:371-384), so the batch is redone on the fp32 CUDA-core kernels
        eng = self.engine(devices[0] if devices else None)
        return self._fp32_engine(eng.device).forward_host(x32, ten_to=ten_to, dtype=out_dtype)

    def predictions_np(self, parameters_dict, devices=None):
        """reference cosmopower_NN.py:388-405 / cosmopower_PCAplusNN.py:384-401 (`devices`: see forward_pass_np)."""
        return self._forward_np(self.dict_to_ordered_arr_np(parameters_dict), ten_to=False, devices=devices)

    def ten_to_predictions_np(self, parameters_dict, devices=None):
        """reference cosmopower_NN.py:409-425 / cosmopower_PCAplusNN.py:405-421 (10**x fused in the kernel)."""
        return self._forward_np(self.dict_to_ordered_arr_np(parameters_dict), ten_to=True, devices=devices)

    # -- batched device-resident twins of predictions_tf / ten_to_predictions_tf ---------------------------
    def predictions_tf(self, parameters_tensor):
        """Array-in/array-out twin (reference cosmopower_NN.py:160-190 / cosmopower_PCAplusNN.py:197-218):
        takes an already ordered CUDA float32 torch tensor (N, n_parameters), returns a CUDA tensor."""
        return self._engine_for(parameters_tensor.shape[0], parameters_tensor.device.index).forward_torch(parameters_tensor, ten_to=False)

    def ten_to_predictions_tf(self, parameters_tensor):
        """reference cosmopower_NN.py:194-211 / cosmopower_PCAplusNN.py:222-239."""
        return self._engine_for(parameters_tensor.shape[0], parameters_tensor.device.index).forward_torch(parameters_tensor, ten_to=True)

    predictions_device = predictions_tf
    ten_to_predictions_device = ten_to_predictions_tf

    def activation(self, x, alpha, beta):
        """Non-linear activation function (reference cosmopower_NN.py:136-156 / cosmopower_PCAplusNN.py:139-158):
        (beta + sigmoid(alpha x) (1 - beta)) x, on NumPy arrays or torch tensors.  (The engine evaluates it inside the
        fused kernel's epilogue; this is the stand-alone helper of the reference API.)"""
        try:
            import torch
            if isinstance(x, torch.Tensor):
                return (beta + torch.sigmoid(alpha * x) * (1.0 - beta)) * x
        except ImportError:
            pass
        with np.errstate(over="ignore"):
            return (beta + (1.0 - beta) / (1.0 + np.exp(-alpha * x))) * x

    def update_emulator_parameters(self):
        """reference cosmopower_NN.py:258-272 / cosmopower_PCAplusNN.py:243-262 copy the TensorFlow variables into the
        NumPy attributes before saving; here the NumPy attributes ARE the parameters, so there is nothing to copy."""
        return None

    # -- out of scope, named so that callers get a clear message ---------------------------------------
    def train(self, *a, **k):
        raise NotImplementedError("cosmopower_b200 is an inference engine; training is out of scope")

    def save(self, filename):
        """Write the flat .npz container (not the pickle) for this model."""
        _loader.save_npz(filename + ".npz", self._attrs(), self._FIELDS)
