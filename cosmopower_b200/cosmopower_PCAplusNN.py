"""Drop-in for the inference path of the reference class `cosmopower_PCAplusNN`
(cosmopower/cosmopower_PCAplusNN.py:14)."""
import numpy as np

from . import _loader
from ._base import _EmulatorBase


class cosmopower_PCAplusNN(_EmulatorBase):
    _FIELDS = _loader.PCANN_FIELDS
    _IS_PCA = True

    def __init__(self, cp_pca=None, n_hidden=[512, 512, 512], restore=False, restore_filename=None,
                 trainable=True, optimizer=None, verbose=False, device=None, precision="f16x3"):
        # signature: reference cosmopower_PCAplusNN.py:37-45 (+ device / precision extensions)
        self._init_runtime(device, precision, verbose)
        if restore is True:
            self.restore(restore_filename)
        else:
            raise NotImplementedError(
                "cosmopower_b200 is inference-only: cosmopower_PCAplusNN needs restore=True "
                "(the cp_pca training branch, reference cosmopower_PCAplusNN.py:59-78, is out of scope)")
        if self.verbose:
            print("\nInitialized cosmopower_PCAplusNN model (B200 engine), \n"
                  "mapping %d input parameters to %d PCA components \n"
                  "and then inverting the PCA compression to obtain %d modes \n"
                  "The model uses %d hidden layers, \nwith %s nodes, respectively. \n"
                  % (self.n_parameters, self.n_pcas, self.n_modes, len(self.n_hidden), list(self.n_hidden)))

    # -- PCA coefficients (reference cosmopower_PCAplusNN.py:163-193) -----------------------------------------
    def _coefficient_attrs(self):
        """The network up to the PCA coefficients is a cosmopower_NN whose output de-standardisation is (pca_std_, pca_mean_):
        `(x.W + b) * pca_std + pca_mean` (reference :190-193) is exactly the NN output layer (cosmopower_NN.py:381-384)."""
        a = {k: getattr(self, k) for k in ("parameters", "n_parameters", "n_hidden", "n_layers", "W_", "b_", "alphas_", "betas_",
                                            "parameters_mean_", "parameters_std_")}
        a["n_modes"] = int(self.n_pcas)
        a["modes"] = np.arange(int(self.n_pcas))
        a["architecture"] = [int(self.n_parameters)] + list(self.n_hidden) + [int(self.n_pcas)]
        a["features_mean_"] = self.pca_mean_
        a["features_std_"] = self.pca_std_
        return a

    def forward_pass_tf(self, parameters_tensor):
        """Forward pass through the network to the PCA coefficients (reference cosmopower_PCAplusNN.py:163-193), as the
        array-in/array-out twin on CUDA float32 torch tensors: (N, n_parameters) -> (N, n_pcas).  Runs on its own engine,
        created on first use, with the coefficient layer as the output layer - on the CUDA-core fp32 kernels: the high-order
        coefficients are small differences of large hidden activations (|z| up to 23 standard deviations for cmb_PP), an
        ill-conditioned quantity that the reference's own float32 pass resolves to 1.5e-4 of a coefficient's spread; the
        spectrum path (predictions_tf) does not go through here."""
        from . import _engine
        dev = parameters_tensor.device.index
        if not hasattr(self, "_coef_engines"):
            self._coef_engines = {}
        if dev not in self._coef_engines:
            self._coef_engines[dev] = _engine.Engine(self._coefficient_attrs(), device=dev, precision="fp32")
        return self._coef_engines[dev].forward_torch(parameters_tensor, ten_to=False)

    def close(self):
        for e in getattr(self, "_coef_engines", {}).values():
            e.close()
        self._coef_engines = {}
        super().close()

    def _validate(self):
        super()._validate()
        if tuple(np.shape(self.pca_transform_matrix_)) != (int(self.n_pcas), int(self.n_modes)):
            raise ValueError("pca_transform_matrix_ must have shape (n_pcas, n_modes)")
