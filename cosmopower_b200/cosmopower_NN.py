"""Drop-in for the inference path of the reference class `cosmopower_NN`
(cosmopower/cosmopower_NN.py:14): same constructor keywords, attribute names and prediction
methods, computed by the B200 engine."""
import numpy as np

from . import _loader
from ._base import _EmulatorBase


class cosmopower_NN(_EmulatorBase):
    _FIELDS = _loader.NN_FIELDS

    def __init__(self, parameters=None, modes=None, parameters_mean=None, parameters_std=None,
                 features_mean=None, features_std=None, n_hidden=[512, 512, 512], restore=False,
                 restore_filename=None, trainable=True, optimizer=None, verbose=False,
                 device=None, precision="f16x3", weights=None):
        # signature: reference cosmopower_NN.py:46-59 (+ device / precision / weights extensions)
        self._init_runtime(device, precision, verbose)
        if restore is True:
            self.restore(restore_filename)
        else:
            # non-restore branch (reference :71-80): attributes from arguments.  The reference then
            # random-initialises TF variables for training; here trained arrays may be supplied via
            # `weights` = dict(W_=[...], b_=[...], alphas_=[...], betas_=[...]).
            self.parameters = parameters
            self.n_parameters = len(self.parameters)
            self.modes = modes
            self.n_modes = len(self.modes)
            self.n_hidden = n_hidden
            self.architecture = [self.n_parameters] + list(self.n_hidden) + [self.n_modes]
            self.n_layers = len(self.architecture) - 1
            one = lambda v, n, fill: np.full(n, fill, np.float32) if v is None else np.asarray(v, np.float32)
            self.parameters_mean_ = one(parameters_mean, self.n_parameters, 0.)
            self.parameters_std_ = one(parameters_std, self.n_parameters, 1.)
            self.features_mean_ = one(features_mean, self.n_modes, 0.)
            self.features_std_ = one(features_std, self.n_modes, 1.)
            if weights is None:
                raise NotImplementedError(
                    "cosmopower_b200 is inference-only: pass restore=True or trained arrays via weights=")
            for k in ("W_", "b_", "alphas_", "betas_"):
                setattr(self, k, [np.asarray(a, np.float32) for a in weights[k]])
            self._validate()
        if self.verbose:
            print("\nInitialized cosmopower_NN model (B200 engine), \n"
                  "mapping %d input parameters to %d output modes, \n"
                  "using %d hidden layers, \nwith %s nodes, respectively. \n"
                  % (self.n_parameters, self.n_modes, len(self.n_hidden), list(self.n_hidden)))

    # -- rescaled predictions (reference cosmopower_NN.py:213-251, 429-466) -----------------------------------
    # The reference's `train()` (:666-714) stores a feature post-processing - closures `postprocessing_np` /
    # `postprocessing_tf` and their vectors - which `rescaled_predictions_*` apply to the network output.  Neither is
    # pickled by `save()`, so a restored reference model raises AttributeError there; here the same four methods exist,
    # raise the same AttributeError until a post-processing is attached, and `set_feature_processing` (extension)
    # attaches the three kinds `train()` knows, with the vectors the caller kept from training.
    def set_feature_processing(self, preprocessing, processing_vectors=None):
        """Attach the feature post-processing of the reference's `train(preprocessing=...)` (cosmopower_NN.py:666-714):
        'mean_sigma' (vectors 'mean', 'sigma'), 'min_max' (vectors 'min', 'max') or None (identity)."""
        if preprocessing == 'mean_sigma':
            vec = {k: np.asarray(processing_vectors[k]) for k in ('mean', 'sigma')}
            post_np = lambda features, v: features * v['sigma'] + v['mean']
            post_tf = lambda features, v: features * v['sigma'] + v['mean']
        elif preprocessing == 'min_max':
            vec = {k: np.asarray(processing_vectors[k]) for k in ('min', 'max')}
            post_np = lambda features, v: features * (v['max'] - v['min']) + v['min']
            post_tf = lambda features, v: features * (v['max'] - v['min']) + v['min']
        elif preprocessing is None:
            vec = {}
            post_np = lambda features, v: features
            post_tf = lambda features, v: features
        else:
            raise ValueError("preprocessing must be 'mean_sigma', 'min_max' or None")
        for k, a in vec.items():
            if a.shape != (int(self.n_modes),):
                raise ValueError("processing vector %r must have shape (%d,)" % (k, self.n_modes))
        self.postprocessing_np = post_np
        self.postprocessing_tf = post_tf
        self.processing_vectors_np = vec
        self._processing_vectors_tf = {}          # per device, built on first use

    @property
    def processing_vectors_tf(self):
        if not hasattr(self, 'processing_vectors_np'):
            raise AttributeError("'cosmopower_NN' object has no attribute 'processing_vectors_tf'")
        return self._processing_vectors_tf

    def _vectors_on(self, device):
        import torch
        v = self._processing_vectors_tf.get(device)
        if v is None:
            v = {k: torch.as_tensor(np.asarray(a, np.float32), device=device) for k, a in self.processing_vectors_np.items()}
            self._processing_vectors_tf[device] = v
        return v

    def rescaled_predictions_np(self, parameters_dict):
        """output predictions * scaling_division + scaling_subtraction (reference cosmopower_NN.py:429-446)."""
        return self.postprocessing_np(self.predictions_np(parameters_dict), self.processing_vectors_np)

    def ten_to_rescaled_predictions_np(self, parameters_dict):
        """reference cosmopower_NN.py:450-466."""
        return 10.**self.rescaled_predictions_np(parameters_dict)

    def rescaled_predictions_tf(self, parameters_tensor):
        """Array-in/array-out twin on CUDA torch tensors (reference cosmopower_NN.py:213-231)."""
        post = self.postprocessing_tf
        return post(self.predictions_tf(parameters_tensor), self._vectors_on(parameters_tensor.device))

    def ten_to_rescaled_predictions_tf(self, parameters_tensor):
        """reference cosmopower_NN.py:235-251."""
        return 10.**self.rescaled_predictions_tf(parameters_tensor)
