"""
pixel_nerf_multiscale_b200 -- B200-native (sm_100a) implementation of pixelNeRF's ray-rendering
hot path behind the Python API of Zxhh123/pixel-nerf-multiscale:

    make_model(conf["model"]) -> PixelNeRFNet   (.encode / .forward / .load_weights)
    NeRFRenderer.from_conf(conf["renderer"], ...).bind_parallel(net, gpus, simple_output)

The arithmetic of the path lives in csrc/ (hand-written CUDA behind the C ABI of
include/pixelnerf_b200.h); this package is the host-side mirror of the reference interface.
"""
from . import util  # noqa: F401
from .model import PixelNeRFNet, make_model  # noqa: F401
from .render import NeRFRenderer  # noqa: F401

__all__ = ["make_model", "PixelNeRFNet", "NeRFRenderer", "util"]
