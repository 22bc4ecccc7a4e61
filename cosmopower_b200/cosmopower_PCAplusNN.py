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

    def _validate(self):
        super()._validate()
        if tuple(np.shape(self.pca_transform_matrix_)) != (int(self.n_pcas), int(self.n_modes)):
            raise ValueError("pca_transform_matrix_ must have shape (n_pcas, n_modes)")
