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
