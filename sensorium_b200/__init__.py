"""sensorium_b200 — B200-native (sm_100a) DwiseNeuro training / inference hot path.

Drop-in surface of lRomul/sensorium for this path: ``DwiseNeuro`` (src/models/dwiseneuro.py),
``MicePoissonLoss`` (src/losses.py), ``MouseModel`` (src/argus_models.py), ``ModelEma`` (src/ema.py),
``Predictor`` (src/predictors.py).  All arithmetic runs in hand-written CUDA kernels behind the C ABI of
``libdwn_b200.so`` (include/dwn_b200.h); there is no CPU or eager-PyTorch fallback.
"""
from .dwiseneuro import DwiseNeuro  # noqa: F401

__all__ = ["DwiseNeuro"]
