"""Drop-in for the reference's `src/model` package: `from model import make_model`."""
from pixel_nerf_multiscale_b200.model import PixelNeRFNet, make_model  # noqa: F401
from pixel_nerf_multiscale_b200.model.code import PositionalEncoding  # noqa: F401
from pixel_nerf_multiscale_b200.model.encoder import ImageEncoder, SpatialEncoder  # noqa: F401
from pixel_nerf_multiscale_b200.model.resnetfc import ResnetFC  # noqa: F401
