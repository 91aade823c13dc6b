"""desmo_b200: B200-native (sm_100a) DESMO training hot path behind the reference's PyTorch surface.

Importing the package never touches the GPU; constructing a model / engine loads ``libdesmo_b200.so`` (built in-tree with
nvcc) and fails loudly when it, or a CUDA device, is missing -- there is no CPU fallback.
"""
from ._lib import DesmoError, PATH_AUTO, PATH_FP32, PATH_GEMM, PATH_TC  # noqa: F401
from .engine import REFERENCE_LRS, DesmoEngine  # noqa: F401
from .model import DESMO, DESMOFourier  # noqa: F401
from .autoencoder import Autoencoder_Linear_Temporal, SINDyAutoencoder  # noqa: F401
from .trainer import DesmoTrainer, PlateauScheduler  # noqa: F401

__all__ = ["DESMO", "DESMOFourier", "SINDyAutoencoder", "Autoencoder_Linear_Temporal", "DesmoEngine", "DesmoTrainer", "PlateauScheduler", "DesmoError", "REFERENCE_LRS",
           "PATH_AUTO", "PATH_FP32", "PATH_TC", "PATH_GEMM"]
