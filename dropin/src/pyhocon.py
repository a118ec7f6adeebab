"""`from pyhocon import ConfigFactory` for drivers that read confs themselves (pyhocon is optional)."""
from pixel_nerf_multiscale_b200.util.conf import ConfigFactory, ConfigTree  # noqa: F401
