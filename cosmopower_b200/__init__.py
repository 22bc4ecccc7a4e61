"""cosmopower_b200 — B200-native engine for CosmoPower's batched emulator forward pass.

Drop-in for `cosmopower.cosmopower_NN` / `cosmopower.cosmopower_PCAplusNN` on the inference path
(restore + predictions_np / ten_to_predictions_np + device-resident batched twins)."""
from .cosmopower_NN import cosmopower_NN
from .cosmopower_PCAplusNN import cosmopower_PCAplusNN
from . import priors
from .fused import predictions_multi_tf, ten_to_sum_predictions_np, ten_to_sum_predictions_tf

__all__ = ["cosmopower_NN", "cosmopower_PCAplusNN", "priors", "predictions_multi_tf", "ten_to_sum_predictions_np",
           "ten_to_sum_predictions_tf"]
__version__ = "0.1.0"
