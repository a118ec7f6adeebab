from .nerf import NeRFRenderer, RenderOutput  # noqa: F401
