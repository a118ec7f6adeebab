"""Drop-in for the reference's `src/render` package: `from render import NeRFRenderer`."""
from pixel_nerf_multiscale_b200.render import NeRFRenderer  # noqa: F401
