"""Likelihoods that consume the emulators on the device (reference: cosmopower/likelihoods/)."""
from .planck2018_lite import tf_planck2018_lite

__all__ = ["tf_planck2018_lite"]
