"""gazenerf_b200 -- B200-native (sm_100a) volumetric renderer behind GazeNeRF's ``GazeNeRFNet.forward`` API.

Only what the hot path needs lives here: ``csrc/`` (CUDA kernels + the C ABI of include/gnrf.h), the ctypes loader,
and the host-side mirror of the reference's module / option / render-util interface.
"""
from .options import BaseOptions  # noqa: F401
from .net import GazeNeRFNet, NeuralRendererParams, RadianceMLP  # noqa: F401
from .render_utils import RenderUtils  # noqa: F401
from .losses import GazeNeRFLoss  # noqa: F401
from .data import SampleStager, sample_to_device  # noqa: F401
from ._lib import build, lib  # noqa: F401

__all__ = ["BaseOptions", "GazeNeRFNet", "NeuralRendererParams", "RadianceMLP", "RenderUtils", "GazeNeRFLoss", "SampleStager", "sample_to_device", "build", "lib"]
